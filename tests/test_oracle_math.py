"""Pins the [IL] quaternion restatement (oracle/il_math.py) with closed-form identities."""
import math

import torch

from oracle import il_math as M


def _rand_quat(n, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    return q / q.norm(dim=1, keepdim=True)


def test_apply_inverse_round_trip():
    q = _rand_quat(256, 0)
    v = torch.randn(256, 3, generator=torch.Generator().manual_seed(1))
    torch.testing.assert_close(M.quat_apply_inverse(q, M.quat_apply(q, v)), v, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(M.quat_apply(q, v).norm(dim=1), v.norm(dim=1), rtol=1e-5, atol=1e-6)


def test_mul_matches_hamilton_product_and_inverse():
    a, b = _rand_quat(128, 2), _rand_quat(128, 3)
    w1, x1, y1, z1 = a.unbind(-1)
    w2, x2, y2, z2 = b.unbind(-1)
    ref = torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                       w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], dim=-1)
    torch.testing.assert_close(M.quat_mul(a, b), ref, rtol=1e-5, atol=1e-6)
    ident = M.quat_mul(M.quat_inv(a), a)
    torch.testing.assert_close(ident, torch.tensor([1.0, 0, 0, 0]).expand_as(ident), rtol=1e-5, atol=1e-6)
    # rotating by a*b == rotating by b then a
    v = torch.randn(128, 3, generator=torch.Generator().manual_seed(4))
    torch.testing.assert_close(M.quat_apply(M.quat_mul(a, b), v), M.quat_apply(a, M.quat_apply(b, v)), rtol=1e-4, atol=1e-5)


def test_euler_round_trip_and_ranges():
    g = torch.Generator().manual_seed(5)
    r = (torch.rand(512, generator=g) - 0.5) * 2 * 3.0
    p = (torch.rand(512, generator=g) - 0.5) * 2 * 1.4
    y = (torch.rand(512, generator=g) - 0.5) * 2 * 3.0
    q = M.quat_from_euler_xyz(r, p, y)
    torch.testing.assert_close(q.norm(dim=1), torch.ones(512), rtol=1e-5, atol=1e-6)
    r2, p2, y2 = M.euler_xyz_from_quat(q)
    for a, b in ((r, r2), (p, p2), (y, y2)):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)
        assert float(b.abs().max()) <= math.pi + 1e-6
    r3, _, y3 = M.euler_xyz_from_quat(q, wrap_to_2pi=True)
    assert float(r3.min()) >= 0.0 and float(y3.max()) < 2 * math.pi + 1e-6


def test_yaw_only_quaternion_is_invariant_to_sign():
    q = _rand_quat(64, 6)
    v = torch.randn(64, 3, generator=torch.Generator().manual_seed(7))
    torch.testing.assert_close(M.quat_apply_inverse(q, v), M.quat_apply_inverse(-q, v), rtol=1e-6, atol=1e-7)
