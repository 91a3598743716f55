"""Engine: eager iteration == CUDA-graph replay == split-graph replay (the multi-process capture layout) on one GPU."""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _params(eng):
    return eng.alg.optimizer.flat.clone()


@pytest.mark.parametrize("split", [False, True])
def test_graph_replay_matches_eager_iteration(cuda, lt_lib, split):
    from locotouch_b200.engine import HotPathEngine

    cfg = dict(num_envs=256, task="teacher", tactile=True, device=cuda, seed=3, num_state_sets=3, hidden=(64, 32), tf32=False)
    torch.manual_seed(0)
    a = HotPathEngine(**cfg)
    b = HotPathEngine(**cfg)
    H.assert_equal(_params(a), _params(b), "identical initial parameters")
    # eager reference: 2 warm-up iterations (capture() runs the same two) + 2 measured
    perms = []
    for it in range(4):
        torch.manual_seed(100 + it)
        a.iteration()
        perms.append(a.perm.clone())
    torch.manual_seed(100)
    b_calls = {"n": 0}
    orig = b.draw_permutation

    def fixed_perm():
        b.perm.copy_(perms[min(b_calls["n"], 3)])
        b_calls["n"] += 1

    b.draw_permutation = fixed_perm
    b.capture(split=split)
    for _ in range(2):
        b.replay()
    torch.cuda.synchronize()
    H.assert_close(_params(b), _params(a), "parameters after 4 iterations (graph replay vs eager)", rtol=1e-4, atol=1e-5)
    H.assert_equal(b.step_counter, a.step_counter, "device step counter")
    ra, rb = a.read_results(), b.read_results()
    assert abs(ra["mean_reward"] - rb["mean_reward"]) <= 1e-5 * max(1.0, abs(ra["mean_reward"]))
    assert abs(ra["learning_rate"] - rb["learning_rate"]) <= 1e-9


def test_prefetched_upload_replay_matches_eager_iteration(cuda, lt_lib):
    """replay(upload=True): the state sets arrive over the copy stream (two device banks, uploads overlapping compute) and
    the result is the one the eager iteration over device-resident sets produces."""
    from locotouch_b200.engine import HotPathEngine

    cfg = dict(num_envs=256, task="teacher", tactile=True, device=cuda, seed=5, num_state_sets=3, hidden=(64, 32), tf32=False)
    a = HotPathEngine(**cfg)
    b = HotPathEngine(**cfg, prefetch=True, pin_host=True)  # 24 % 3 == 0: both banks cycle the same host sets as `a`
    assert len(b.dev_flat) == 2 * b.T and len(b.host_flat) == 3
    perms = []
    for it in range(5):
        torch.manual_seed(200 + it)
        a.iteration()
        perms.append(a.perm.clone())
    calls = {"n": 0}

    def fixed_perm():
        b.perm.copy_(perms[min(calls["n"], 4)])
        calls["n"] += 1

    b.draw_permutation = fixed_perm
    b.capture(split=False)
    # scribble over the device banks: replay(upload=True) must restore every set from the host before it is read
    for flat in b.dev_flat:
        flat.fill_(0x7f)
    for _ in range(3):
        b.replay(upload=True)
    torch.cuda.synchronize()
    H.assert_close(_params(b), _params(a), "parameters after 5 iterations (prefetched uploads vs eager)", rtol=1e-4, atol=1e-5)
    H.assert_equal(b.step_counter, a.step_counter, "device step counter")


def test_fused_action_term_matches_separate_launch(cuda, lt_lib):
    """K0 inside K1 (``FusedMdp.step(actions=...)``): the action-term state, rewards, dones and both observation groups of a rollout are
    bit-identical to ``lt_process_actions`` followed by ``lt_mdp_step`` (reference mdp/actions.py:30-44 runs before the managers)."""
    from locotouch_b200.engine import HotPathEngine

    cfg = dict(num_envs=300, task="teacher", tactile=True, device=cuda, seed=11, num_state_sets=3, hidden=(64, 32), tf32=False)
    a = HotPathEngine(**cfg)
    b = HotPathEngine(**cfg)
    a.fuse_action_term, b.fuse_action_term = True, False
    for eng in (a, b):
        torch.manual_seed(7)
        eng.rollout()
    torch.cuda.synchronize()
    sa, sb = a.alg.storage, b.alg.storage
    H.assert_equal(sa.actions, sb.actions, "actions (same policy, same draws)")
    for name in ("raw_actions", "prev_raw_actions", "prev_prev_raw_actions", "processed_actions"):
        H.assert_equal(getattr(a.action_term, name), getattr(b.action_term, name), name)
    H.assert_equal(sa.rewards, sb.rewards, "rewards")
    H.assert_equal(sa.dones, sb.dones, "dones")
    H.assert_equal(sa._obs_buf, sb._obs_buf, "policy observations")
    H.assert_equal(sa._priv_buf, sb._priv_buf, "critic observations")


def test_fused_rollout_store_matches_separate_launch(cuda, lt_lib):
    """K3 inside K1 (``FusedMdp.step(store=...)``): the bootstrapped rewards and the dones of every rollout row are bit-identical to
    ``lt_store_step`` after ``lt_mdp_step`` (reference ppo.py:162-165 + rollout_storage.py:86-88), time-outs included, at a ragged env
    count; and the step needs one launch less."""
    from locotouch_b200 import _C
    from locotouch_b200.engine import HotPathEngine

    cfg = dict(num_envs=301, task="teacher", tactile=True, device=cuda, seed=13, num_state_sets=3, hidden=(64, 32), tf32=False)
    a = HotPathEngine(**cfg)
    b = HotPathEngine(**cfg)
    a.fuse_store, b.fuse_store = True, False
    launches = []
    for eng in (a, b):
        torch.manual_seed(9)
        n0 = _C.launch_count
        eng.rollout()
        launches.append(_C.launch_count - n0)
    torch.cuda.synchronize()
    sa, sb = a.alg.storage, b.alg.storage
    assert int(sb.dones.sum()) > 0, "the rollout must contain resets"
    H.assert_equal(sa.actions, sb.actions, "actions (same policy, same draws)")
    H.assert_equal(sa.rewards, sb.rewards, "stored (bootstrapped) rewards")
    H.assert_equal(sa.dones, sb.dones, "stored dones")
    H.assert_equal(sa.values, sb.values, "values")
    H.assert_equal(sa.returns, sb.returns, "returns")
    H.assert_equal(sa.advantages, sb.advantages, "advantages")
    H.assert_equal(sa._obs_buf, sb._obs_buf, "policy observations")
    assert launches[1] - launches[0] == a.T, launches
