"""K2 parity: binary taxel synthesis + delay line vs the reference golden bitmaps and the oracle (bit-exact)."""
import pytest
import torch

from oracle import tactile as OT
from tests import helpers as H
from tests.golden.make_golden import TACTILE, tactile_inputs
from locotouch_b200.sim import synth

pytestmark = pytest.mark.gpu

CASES = [(0.005, 0.005, 0.0), (0.2, 0.1, 0.3)]


@pytest.mark.parametrize("case", [0, 1])
def test_bitmaps_and_delay_line_match_reference(cuda, lt_lib, case):
    from locotouch_b200 import ops

    gold = H.load_golden("tactile_c4.npz")
    c = TACTILE
    n = c["N"]
    p_drop, p_add, jitter = CASES[case]
    env = synth.make_env(n, seed=c["seed"] + case, with_object=True, with_tactile=True, tactile_jitter=jitter)
    thr = torch.as_tensor(gold[f"thresholds_{case}"]).reshape(n, 221).to(cuda)
    # fused packed delay line (28 B/env/frame) and the generic fp32 one (drop-in TactileRecorder)
    ring_p = torch.zeros(n, 2, 7, device=cuda, dtype=torch.int32)
    first_p = torch.ones(n, device=cuda, dtype=torch.uint8)
    ring_f = torch.zeros(n, 2, 442, device=cuda)
    first_f = torch.ones(n, device=cuda, dtype=torch.uint8)
    delay = torch.ones(n, device=cuda, dtype=torch.int64)
    for step in range(c["steps"]):
        u_drop, u_add = tactile_inputs(step, n, case)
        q = env.scene["robot"].data.body_quat_w.to(cuda)
        f = env.scene.sensors["tactile_contact_sensor"].data.net_forces_w.to(cuda)
        if step == 2:
            ids = torch.arange(0, n, 3, device=cuda)
            ring_p[ids] = 0
            first_p[ids] = 1
            ring_f[ids] = 0
            first_f[ids] = 1
        delayed = torch.empty(n, 442, device=cuda)
        fn = torch.empty(n, 221, device=cuda)
        orig = torch.empty(n, 221, device=cuda, dtype=torch.uint8)
        sig, packed = ops.taxel_synth(q, f, thr, quat_body_offset=17, u_drop=u_drop.reshape(n, 221).to(cuda), u_add=u_add.reshape(n, 221).to(cuda),
                                      p_drop=p_drop, p_add=p_add, normal_forces=fn, original_contact=orig, delay_ring=ring_p,
                                      delay_first=first_p, delay_steps=delay, delayed_signal=delayed)
        torch.cuda.synchronize()
        H.assert_equal(sig, gold[f"signal_{case}"][step], f"bitmap step {step} vs reference")
        H.assert_equal(delayed, gold[f"delayed_{case}"][step], f"fused delay line step {step} vs reference")
        out_f = ops.tactile_delay(ring_f, first_f, delay, sig)
        H.assert_equal(out_f, gold[f"delayed_{case}"][step], f"fp32 delay line step {step} vs reference")
        # oracle on the same inputs: side buffers and the packed words
        res = OT.binary_taxels(q[:, 17:].cpu(), f.cpu(), thr.cpu(), u_drop.reshape(n, 221), u_add.reshape(n, 221), p_drop, p_add)
        H.assert_equal(orig, res["original"], "original_contact_taxels")
        H.assert_equal(packed, OT.pack_bits(res["contact"]), "ballot-packed words")
        H.assert_close(fn, res["normal_forces"], "original_normal_forces", rtol=1e-6, atol=1e-7)
        synth.advance(env, tactile_jitter=jitter)
    assert int(first_p.sum()) == 0 and int(first_f.sum()) == 0


def test_threshold_boundary_is_strict(cuda, lt_lib):
    """F_n == thr is NOT a contact (strict '>' at reference observations.py:159); identity quaternion makes F_n exact."""
    from locotouch_b200 import ops

    n, t = 5, 221
    q = torch.zeros(n, t, 4, device=cuda)
    q[..., 0] = 1.0
    g = torch.Generator().manual_seed(0)
    fz = torch.rand(n, t, generator=g)
    f = torch.zeros(n, t, 3)
    f[..., 2] = -fz
    thr = fz.clone()
    thr[:, ::2] = torch.nextafter(fz[:, ::2], torch.zeros(()))  # just below the force on even taxels -> contact
    sig, _ = ops.taxel_synth(q, f.to(cuda), thr.to(cuda), p_drop=0.0, p_add=0.0)
    expect = torch.zeros(n, t)
    expect[:, ::2] = 1.0
    H.assert_equal(sig[:, :t], expect, "strict threshold")
    H.assert_equal(sig[:, t:], expect, "second channel identical")


@pytest.mark.parametrize("n", [1, 405, 4097])
def test_taxels_match_oracle_on_ragged_sizes_and_rng_path(cuda, lt_lib, n):
    from locotouch_b200 import ops

    env = synth.make_env(n, seed=n, with_object=True, with_tactile=True, tactile_jitter=0.2)
    g = torch.Generator().manual_seed(n)
    thr = 0.05 + (torch.rand(n, 221, generator=g) - 0.5) * 0.02
    ud, ua = torch.rand(n, 221, generator=g), torch.rand(n, 221, generator=g)
    q = env.scene["robot"].data.body_quat_w
    f = env.scene.sensors["tactile_contact_sensor"].data.net_forces_w
    res = OT.binary_taxels(q[:, 17:], f, thr, ud, ua, 0.05, 0.05)
    sig, packed = ops.taxel_synth(q.to(cuda), f.to(cuda), thr.to(cuda), quat_body_offset=17, u_drop=ud.to(cuda), u_add=ua.to(cuda), p_drop=0.05, p_add=0.05)
    H.assert_equal(sig, res["signal"], "bitmap")
    H.assert_equal(packed, OT.pack_bits(res["contact"]), "packed")
    if n >= 405:
        # production mode: in-kernel Philox uniforms -> dropout / addition rates match the probabilities
        s0, _ = ops.taxel_synth(q.to(cuda), f.to(cuda), thr.to(cuda), quat_body_offset=17, p_drop=0.0, p_add=0.0)
        s1, _ = ops.taxel_synth(q.to(cuda), f.to(cuda), thr.to(cuda), quat_body_offset=17, p_drop=0.1, p_add=0.05, seed=3, offset=1)
        base, noisy = s0[:, :221].bool(), s1[:, :221].bool()
        added = (noisy & ~base).float().sum() / (~base).float().sum()
        dropped = (~noisy & base).float().sum() / base.float().sum()
        assert abs(float(added) - 0.05) < 0.01, float(added)
        assert abs(float(dropped) - 0.1 * 0.95) < 0.02, float(dropped)  # a dropped taxel can be re-added
        s2, _ = ops.taxel_synth(q.to(cuda), f.to(cuda), thr.to(cuda), quat_body_offset=17, p_drop=0.1, p_add=0.05, seed=3, offset=1)
        H.assert_equal(s2, s1, "counter RNG reproducible")
