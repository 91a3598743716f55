"""Host-side logic that needs no GPU: the fork / join helper without a CUDA stream, RolloutStorage's "already in its row" rule
(K3 inside K1: nothing to launch), and the argument validation of the fused action-term / rollout-store fields of lt_mdp_step."""
import ctypes

import pytest
import torch


def test_side_stream_degrades_to_inline_execution_without_cuda():
    from locotouch_b200.streams import SideStream

    s = SideStream("cpu")
    assert s.stream is None and s.mark() is None
    ran = []
    with s.forked():
        ran.append(1)
    with s.forked(after=s.mark()):
        ran.append(2)
    s.join()
    assert ran == [1, 2]


def _storage(T=3, N=5, D=4, A=2):
    from locotouch_b200.loco_rl.storage.rollout_storage import RolloutStorage

    return RolloutStorage(N, T, (D,), (D,), (A,), device="cpu")


def _transition(st, s, in_place: bool):
    t = st.Transition()
    t.observations, t.critic_observations = st._obs_buf[s], st._priv_buf[s]
    t.actions, t.values, t.actions_log_prob = st.actions[s], st.values[s], st.actions_log_prob[s]
    t.action_mean, t.action_sigma = st.mu[s], st.sigma[s]
    if in_place:
        t.rewards, t.dones = st.rewards[s].view(-1), st.dones[s].view(-1)
    else:
        t.rewards, t.dones = torch.zeros(st.num_envs), torch.zeros(st.num_envs, dtype=torch.uint8)
    return t


def test_add_transitions_launches_nothing_when_everything_is_in_its_row():
    """Rewards (already bootstrapped) and dones written into the rollout row by the MDP launch, observations built in place: the
    store is a no-op -- it runs on CPU tensors, where any launch would raise (the library has no CPU path)."""
    st = _storage()
    st.rewards[0] += 1.5
    st.add_transitions(_transition(st, 0, in_place=True))
    assert st.step == 1 and float(st.rewards[0].sum()) == 1.5 * st.num_envs
    st.add_transitions(_transition(st, 1, in_place=True))
    assert st.step == 2


def test_add_transitions_still_stores_everything_else():
    """Not in place (separate reward tensor, or a time-out mask to bootstrap with): the K3 launch is needed -> on CPU tensors it raises."""
    from locotouch_b200._C import LocoTouchLibraryError

    st = _storage()
    with pytest.raises(LocoTouchLibraryError):
        st.add_transitions(_transition(st, 0, in_place=False))
    st = _storage()
    with pytest.raises(LocoTouchLibraryError):  # in-place tensors but a time-out mask: the bootstrap has not been applied yet
        st.add_transitions(_transition(st, 0, in_place=True), time_outs=torch.zeros(st.num_envs, dtype=torch.bool), gamma=0.99)
    st = _storage()
    t = _transition(st, 1, in_place=True)  # row 1's tensors while the storage is at step 0: not the current row
    with pytest.raises(LocoTouchLibraryError):
        st.add_transitions(t)


def test_overflow_is_reported_like_the_reference():
    st = _storage(T=1)
    st.add_transitions(_transition(st, 0, in_place=True))
    with pytest.raises(OverflowError):
        st.add_transitions(_transition(st, 0, in_place=True))


def test_mdp_step_rejects_fused_fields_without_the_reward_phase(lt_lib):
    """act_new / store_rewards belong to the reward pass (validated before any CUDA call)."""
    from locotouch_b200 import _C

    a = _C.LtMdpArgs()
    a.N, a.J = 8, 12
    a.phases = _C.LT_PHASE_OBS
    dummy = ctypes.c_float(0.0)
    a.act_new = ctypes.addressof(dummy)
    assert lt_lib.lt_mdp_step(ctypes.byref(a), None) == 1
    a.act_new = None
    a.store_rewards = ctypes.addressof(dummy)
    assert lt_lib.lt_mdp_step(ctypes.byref(a), None) == 1
    assert a.act_reset_on_done == 0 and a.store_gamma == 0.0
