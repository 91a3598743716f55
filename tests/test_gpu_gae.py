"""K4 parity: lt_gae (through the C ABI) vs the oracle and the reference golden vectors."""
import pytest
import torch

from oracle import ppo as OP
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _run(r, cuda, normalize=True, gamma=0.99, lam=0.95):
    from locotouch_b200 import ops

    T, N = r["rewards"].shape[:2]
    ret, adv = ops.gae(r["rewards"].to(cuda), r["values"].to(cuda), r["dones"].byte().to(cuda), r["last_values"].to(cuda), gamma, lam, normalize)
    torch.cuda.synchronize()
    return ret.cpu(), adv.cpu()


def test_gae_matches_reference_golden(cuda, lt_lib):
    gold = H.load_golden("ppo_c1.npz")
    r = H.make_rollout(T=24, N=64, seed=0)
    ret, adv = _run(r, cuda, True)
    H.assert_equal(ret, gold["gae_returns"], "returns (bit-exact: same fp32 expression order, no FMA contraction)")
    H.assert_close(adv, gold["gae_advantages"], "normalised advantages")
    _, adv_raw = _run(r, cuda, False)
    H.assert_equal(adv_raw, gold["gae_advantages_raw"], "raw advantages")


@pytest.mark.parametrize("T,N", [(24, 1), (24, 4097), (7, 33), (1, 5), (48, 405)])
def test_gae_matches_oracle_on_ragged_sizes(cuda, lt_lib, T, N):
    r = H.make_rollout(T=T, N=N, seed=T * 1000 + N) if N > 1 else H.make_rollout(T=T, N=2, seed=3)
    if N == 1:
        r = {k: (v[:, :1] if v.dim() == 3 else v[:1]) for k, v in r.items()}
    for normalize in (False,) if T * N < 3 else (True, False):
        ret, adv = _run(r, cuda, normalize)
        o_ret, o_adv = OP.gae_returns(r["rewards"], r["values"], r["dones"].byte(), r["last_values"], 0.99, 0.95, normalize)
        H.assert_equal(ret, o_ret, f"returns T={T} N={N}")
        if normalize:
            H.assert_close(adv, o_adv, f"advantages T={T} N={N}", rtol=1e-5, atol=1e-5)
        else:
            H.assert_equal(adv, o_adv, f"raw advantages T={T} N={N}")


def test_gae_full_size_properties(cuda, lt_lib):
    """BASELINE size (4096 envs x 24 steps): size-independent properties + oracle on the same inputs."""
    r = H.make_rollout(T=24, N=4096, seed=9)
    ret, adv_raw = _run(r, cuda, False)
    H.assert_equal(adv_raw, ret - r["values"], "adv == returns - values")
    # all-done env: no bootstrap -> return == reward ; never-done env at last step: return == r + gamma * V_last
    H.assert_close(ret[:, 0], r["rewards"][:, 0], "all-done env returns its rewards", rtol=1e-6, atol=1e-7)
    _, adv = _run(r, cuda, True)
    assert abs(float(adv.double().mean())) < 1e-6
    assert abs(float(adv.double().std()) - 1.0) < 1e-5
    o_ret, o_adv = OP.gae_returns(r["rewards"], r["values"], r["dones"].byte(), r["last_values"], 0.99, 0.95, True)
    H.assert_equal(ret, o_ret, "returns at full size")
    H.assert_close(adv, o_adv, "advantages at full size", rtol=1e-5, atol=1e-5)
    # linearity in the rewards (dones fixed): GAE(r1 + r2, V1 + V2) == GAE(r1, V1) + GAE(r2, V2) up to rounding
    r2 = H.make_rollout(T=24, N=4096, seed=10)
    r2["dones"] = r["dones"]
    ret2, _ = _run(r2, cuda, False)
    rs = dict(r)
    rs["rewards"], rs["values"], rs["last_values"] = r["rewards"] + r2["rewards"], r["values"] + r2["values"], r["last_values"] + r2["last_values"]
    ret_s, _ = _run(rs, cuda, False)
    H.assert_close(ret_s, ret + ret2, "linearity", rtol=1e-4, atol=1e-5)


def test_gae_rejects_bad_arguments(cuda, lt_lib):
    from locotouch_b200 import _C, ops

    r = H.make_rollout(T=4, N=8, seed=1)
    with pytest.raises(_C.LocoTouchLibraryError):
        ops.gae(r["rewards"].to(cuda), r["values"].to(cuda), r["dones"].to(cuda).float(), r["last_values"].to(cuda), 0.99, 0.95)
