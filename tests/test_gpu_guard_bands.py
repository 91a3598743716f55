"""Own-made write check (compute-sanitizer is closed on this pool, profiles/r2_sanitizer_closed.md): the outputs of the kernels are
carved out of larger allocations whose guard bands hold a sentinel; a launch must leave the bands untouched and overwrite every
element inside."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SENTINEL = -12345.678


class Guarded:
    def __init__(self, device):
        self.device, self.items = device, []

    def make(self, *shape, dtype=torch.float32, band=4096, fill_inside=True):
        n = 1
        for s in shape:
            n *= s
        pad = band // torch.empty(0, dtype=dtype).element_size()
        raw = torch.empty(pad + n + pad, dtype=dtype, device=self.device)
        raw.fill_(SENTINEL if dtype.is_floating_point else 0x5A)
        view = raw[pad:pad + n].view(*shape)
        self.items.append((raw, pad, n, dtype, fill_inside))
        return view

    def check(self, what):
        for raw, pad, n, dtype, inside in self.items:
            s = SENTINEL if dtype.is_floating_point else 0x5A
            assert bool((raw[:pad] == s).all()) and bool((raw[pad + n:] == s).all()), f"{what}: a guard band was written"
            if inside and dtype.is_floating_point:
                assert not bool((raw[pad:pad + n] == s).any()), f"{what}: an output element was never written"


@pytest.mark.parametrize("B,n,k", [(1000, 132, 100), (4097, 512, 348), (333, 12, 128), (77, 1, 128)])
def test_wgrad_writes_only_its_outputs(cuda, lt_lib, B, n, k):
    from locotouch_b200 import ops

    g = Guarded(cuda)
    out, db = g.make(n, k), g.make(n)
    assert ops.wgrad(torch.randn(B, n, device=cuda), torch.randn(B, k, device=cuda), out, db) is not None
    torch.cuda.synchronize()
    g.check("K15")


@pytest.mark.parametrize("B", [1, 405, 4097])
def test_heads_loss_writes_only_its_outputs(cuda, lt_lib, B):
    from locotouch_b200 import ops

    A, H = 12, 128
    g = Guarded(cuda)
    bufs = ops.PpoLossBuffers(B, A, cuda)
    bufs.grad_mu, bufs.grad_value, bufs.grad_sigma = g.make(B, A), g.make(B), g.make(A)
    gha, ghc, mu, val = g.make(B, H), g.make(B, H), g.make(B, A), g.make(B)
    rn = lambda *s: torch.randn(*s, device=cuda)  # noqa: E731
    ops.ppo_heads_loss(rn(B, H), rn(B, H), rn(A, H) * 0.1, rn(A), rn(1, H) * 0.1, rn(1), 0.5 + torch.rand(A, device=cuda), rn(B, A), rn(B), rn(B, A),
                       0.5 + torch.rand(B, A, device=cuda), rn(B), rn(B), rn(B), gha, ghc, buffers=bufs, mu_out=mu, value_out=val)
    torch.cuda.synchronize()
    g.check("K16")


@pytest.mark.parametrize("M", [1, 405, 4097])
def test_student_cnn_and_contact_sensor_write_only_their_outputs(cuda, lt_lib, M):
    from locotouch_b200 import ops

    g = Guarded(cuda)
    rn = lambda *s: torch.randn(*s, device=cuda)  # noqa: E731
    w = (rn(24, 2, 4, 4), rn(24), rn(24, 24, 3, 3), rn(24), rn(24, 24, 2, 2), rn(24), rn(64, 192), rn(64))
    ops.student_cnn_forward(w, image=(torch.rand(M, 442, device=cuda) < 0.1).float(), out=g.make(M, 64))
    hist = g.make(M, 3, 17, 3)
    tm = [g.make(M, 17) for _ in range(4)]
    for t in tm + [hist]:
        t.zero_()
    ops.contact_sensor_update(rn(M, 17, 3), net_forces_w=g.make(M, 17, 3), history=hist, current_air_time=tm[0], last_air_time=tm[1],
                              current_contact_time=tm[2], last_contact_time=tm[3], dt=0.02, reset_mask=(torch.rand(M, device=cuda) < 0.2).byte())
    torch.cuda.synchronize()
    g.check("K17 / K18")


@pytest.mark.parametrize("n", [1, 405, 4097])
def test_gae_and_taxel_write_only_their_outputs(cuda, lt_lib, n):
    from locotouch_b200 import ops

    g = Guarded(cuda)
    T = 24
    ret, adv = g.make(T, n), g.make(T, n)
    ops.gae(torch.randn(T, n, device=cuda), torch.randn(T, n, device=cuda), (torch.rand(T, n, device=cuda) < 0.05).byte(), torch.randn(n, device=cuda),
            0.99, 0.95, n > 1, ret, adv)
    q = torch.randn(n, 238, 4, device=cuda)
    q = q / q.norm(dim=-1, keepdim=True)
    sig = g.make(n, 442)
    packed = g.make(n, 7, dtype=torch.int32)
    ops.taxel_synth(q, torch.randn(n, 221, 3, device=cuda) * 0.1, torch.full((n, 221), 0.05, device=cuda), quat_body_offset=17, seed=1, offset=0,
                    signal=sig, packed=packed)
    torch.cuda.synchronize()
    g.check("K4 / K2")
