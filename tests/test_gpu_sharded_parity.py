"""Env-sharded training == one learner on the concatenated envs, on the CUDA path, on ONE GPU (SURVEY.md section 4 "Multi-GPU" row,
section 8e parity statement; reference loco_rl/loco_rl/algorithms/ppo.py:264-281,350-353).

Two PPO learners (replicas) each own half of a fixed rollout; per mini-batch their flat gradient buffers are summed by
``lt_peer_sum_clip_adam`` with world = 2 over two ordinary device buffers (the ABI takes raw pointers; between processes the same
pointers come from NVLink-mapped symmetric memory), the KL statistic rides in the buffers' tail, the advantage statistics are summed
like the 3-double all-reduce.  Compared with ONE learner that holds all envs and whose mini-batch i is the union of the two shards'
mini-batch i: advantages, the learning-rate sequence of all 20 mini-batch steps, the parameters after the update, and that the two
replicas stay bit-identical."""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu

CFG = dict(num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0, entropy_coef=0.01,
           learning_rate=1.0e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01)
OBS, A, HID, T = 48, 12, [64, 48, 32], 24


def _learner(n, dev):
    from locotouch_b200.loco_rl import PPO, ActorCritic

    torch.manual_seed(0)
    ac = ActorCritic(OBS, OBS, A, HID, HID, "elu", 1.0)
    alg = PPO(ac, device=str(dev), **CFG)
    alg.init_storage(n, T, [OBS], [OBS], [A])
    return alg


def _collect(alg, r, eps, envs, dev):
    ac = alg.actor_critic
    for t in range(T):
        ac.rng = lambda mean, _e=eps[t][envs].contiguous(): _e
        alg.act(r["obs"][t][envs].to(dev), r["critic_obs"][t][envs].to(dev))
        alg.process_env_step(r["rewards"][t, envs, 0].to(dev), r["dones"][t, envs, 0].long().to(dev), {"time_outs": r["time_outs"][t, envs, 0].to(dev)})


def _drive(learners, perms, kl_trace=None):
    """20 mini-batch steps in lock step; returns the learning rate after every step (learner 0's -- all are asserted equal)."""
    lrs = []
    for alg, perm in zip(learners, perms):
        alg.optimizer.sync_lr_to_device()
        alg.update_begin(perm)
    for _epoch in range(CFG["num_learning_epochs"]):
        for i in range(CFG["num_mini_batches"]):
            for alg in learners:
                alg.minibatch_grads(i)
            if kl_trace is not None:
                kl_trace.append([float(alg._loss_bufs.out[4]) for alg in learners])
            for alg in learners:
                alg.allreduce_grads()
            for alg in learners:
                alg.step_after_reduce()
            for alg in learners:
                alg.after_step_barrier()
            now = [float(alg.optimizer.lr_t) for alg in learners]
            assert len(set(now)) == 1, f"replicas took different learning-rate decisions: {now}"
            lrs.append(now[0])
    for alg in learners:
        alg.storage.clear()
    return lrs


@pytest.mark.parametrize("tf32", [False, True])
def test_two_shards_equal_one_learner_on_the_concatenated_rollout(cuda, lt_lib, tf32):
    from locotouch_b200.loco_rl import PPO

    n = 96
    torch.backends.cuda.matmul.allow_tf32 = tf32
    try:
        r = H.make_rollout(T=T, N=2 * n, obs_dim=OBS, A=A, seed=21)
        g = torch.Generator().manual_seed(5)
        eps = torch.randn(T, 2 * n, A, generator=g).to(cuda)
        one = _learner(2 * n, cuda)
        a, b = _learner(n, cuda), _learner(n, cuda)
        PPO.attach_local_peers([a, b])
        assert a.peer_gradients and b.peer_gradients and a._world_info() == (0, 2) and b._world_info() == (1, 2)
        ea, eb = torch.arange(0, n), torch.arange(n, 2 * n)
        _collect(one, r, eps, torch.arange(2 * n), cuda)
        _collect(a, r, eps, ea, cuda)
        _collect(b, r, eps, eb, cuda)
        last = r["critic_obs"][-1].to(cuda)
        one.compute_returns(last)
        assert a.compute_returns_scan(last[ea.to(cuda)]) and b.compute_returns_scan(last[eb.to(cuda)])
        PPO.sum_local_adv_stats([a, b])
        a.compute_returns_normalize()
        b.compute_returns_normalize()
        # ---- advantage mean / std are those of the concatenated envs
        adv_shards = torch.cat([a.storage.advantages, b.storage.advantages], dim=1)
        H.assert_equal(torch.cat([a.storage.returns, b.storage.returns], dim=1), one.storage.returns, "returns")
        H.assert_close(adv_shards, one.storage.advantages, "globally normalised advantages", rtol=1e-6, atol=1e-6)
        # ---- permutations: shard mini-batch i of both learners together == mini-batch i of the single learner
        gp = torch.Generator().manual_seed(9)
        pa, pb = torch.randperm(T * n, generator=gp), torch.randperm(T * n, generator=gp)
        mb = T * n // CFG["num_mini_batches"]
        to_one = lambda idx, first: (idx // n) * (2 * n) + first + idx % n  # [t, env_local] of a shard -> [t, env] of the whole
        p_one = torch.cat([torch.cat([to_one(pa[i * mb:(i + 1) * mb], 0), to_one(pb[i * mb:(i + 1) * mb], n)]) for i in range(CFG["num_mini_batches"])])
        kl_one, kl_sh = [], []
        lr_one = _drive([one], [p_one.to(cuda)], kl_one)
        lr_sh = _drive([a, b], [pa.to(cuda), pb.to(cuda)], kl_sh)
        # the KL the decision sees: mean over the shards' means == the single learner's mean
        kl_mean = torch.tensor([sum(k) / 2 for k in kl_sh])
        H.assert_close(kl_mean, torch.tensor([k[0] for k in kl_one]), "KL statistic of every mini-batch step", rtol=(2e-2 if tf32 else 2e-3), atol=1e-7)
        assert lr_sh == lr_one, f"learning-rate sequences differ:\n sharded {lr_sh}\n single  {lr_one}"
        assert len(set(lr_one)) > 2, "the scenario must exercise the adaptive schedule"
        H.assert_equal(a.optimizer.flat, b.optimizer.flat, "replicas after 20 steps")
        H.assert_equal(a.optimizer.exp_avg_sq, b.optimizer.exp_avg_sq, "Adam state of the replicas")
        # 20 Adam steps; the two sides sum the mini-batch in a different order (two half sums vs one), hence not bit-exact.
        # fp32: parameters agree to <= 1e-5 relative to the parameter scale; TF32 (the production GEMM mode) rounds operands to 10 bits
        H.assert_close(a.optimizer.flat, one.optimizer.flat, "parameters: 2 shards vs 1 learner", rtol=(2e-3 if tf32 else 1e-5), atol=(2e-3 if tf32 else 2e-5))
        la, lo = a.update_epilogue(), one.update_epilogue()
        lb = b.update_epilogue()
        for k in range(3):  # logged means: mean of the shards' means
            assert abs((la[k] + lb[k]) / 2 - lo[k]) <= (5e-3 if tf32 else 1e-4) * max(1.0, abs(lo[k]))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = False


def test_module_to_after_construction_keeps_optimizer_attached(cuda, lt_lib):
    """``actor_critic.to(device)`` after PPO construction (e.g. OnPolicyRunner.get_inference_policy) must not detach the
    fused optimizer from the live parameters."""
    n = 32
    alg = _learner(n, cuda)
    flat0 = alg.optimizer.flat
    alg.actor_critic.to(cuda)            # same device: the flat views survive
    assert alg.actor_critic.flat_params is flat0
    r = H.make_rollout(T=T, N=n, obs_dim=OBS, A=A, seed=3)
    eps = torch.randn(T, n, A, generator=torch.Generator().manual_seed(1)).to(cuda)
    _collect(alg, r, eps, torch.arange(n), cuda)
    alg.compute_returns(r["critic_obs"][-1].to(cuda))
    before = torch.cat([p.detach().flatten().clone() for p in alg.actor_critic.parameters()])
    alg.actor_critic.double().float()    # a real move: every parameter tensor is re-created
    assert alg.actor_critic.flat_params is None
    alg.update()
    after = torch.cat([p.detach().flatten() for p in alg.actor_critic.parameters()])
    assert not torch.equal(before, after), "update() did not change the live parameters"
    assert alg.optimizer.flat is alg.actor_critic.flat_params
    assert all(p.data_ptr() >= alg.optimizer.flat.data_ptr() for p in alg.actor_critic.parameters())
