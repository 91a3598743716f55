"""Drop-in loco_rl classes on the GPU vs one full iteration of the UNMODIFIED reference (golden ppo_c1.npz):
PPO.act x T -> process_env_step -> compute_returns -> update (2 epochs x 2 mini-batches, adaptive learning rate)."""
import pytest
import torch

from tests import helpers as H
from tests import scenarios as S

pytestmark = pytest.mark.gpu


def _build(cuda, gold):
    from locotouch_b200.loco_rl import PPO, ActorCritic

    c = S.PPO_SMALL
    torch.backends.cuda.matmul.allow_tf32 = False  # fp32 GEMMs for the parity run (the bench uses TF32 like the reference)
    ac = ActorCritic(c["obs_dim"], c["obs_dim"], c["A"], c["hidden"], c["hidden"], "elu", 1.0)
    assert [k for k, _ in ac.named_parameters()] == [str(n) for n in gold["ppo_param_names"]]
    flat = torch.as_tensor(gold["ppo_init_params"])
    off = 0
    with torch.no_grad():
        for p in ac.parameters():
            p.copy_(flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
    alg = PPO(ac, num_learning_epochs=2, num_mini_batches=2, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0, entropy_coef=0.01,
              learning_rate=1.0e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cuda:0")
    alg.init_storage(c["N"], c["T"], [c["obs_dim"]], [c["obs_dim"]], [c["A"]])
    return alg, ac


def test_rollout_gae_and_update_match_reference(cuda, lt_lib):
    gold = H.load_golden("ppo_c1.npz")
    c = S.PPO_SMALL
    alg, ac = _build(cuda, gold)
    r = H.make_rollout(T=c["T"], N=c["N"], obs_dim=c["obs_dim"], A=c["A"], seed=c["seed"])
    eps = torch.as_tensor(gold["ppo_eps"]).to(cuda)
    for t in range(c["T"]):
        ac.rng = lambda mean, _e=eps[t]: _e
        obs, cobs = r["obs"][t].to(cuda), r["critic_obs"][t].to(cuda)
        actions = alg.act(obs, cobs)
        assert actions.data_ptr() == alg.storage.actions[t].data_ptr(), "act() must write straight into the rollout slot"
        alg.process_env_step(r["rewards"][t, :, 0].to(cuda), r["dones"][t, :, 0].long().to(cuda), {"time_outs": r["time_outs"][t, :, 0].to(cuda)})
    with pytest.raises(OverflowError):
        alg.storage.add_transitions(alg.transition)
    alg.compute_returns(r["critic_obs"][-1].to(cuda))
    st = alg.storage
    H.assert_close(st.actions, gold["ppo_actions"], "actions", rtol=1e-5, atol=1e-5)
    H.assert_close(st.actions_log_prob, gold["ppo_logp"], "log prob", rtol=1e-5, atol=1e-4)
    H.assert_close(st.values, gold["ppo_values"], "values", rtol=1e-5, atol=1e-5)
    H.assert_close(st.rewards, gold["ppo_rewards"], "bootstrapped rewards", rtol=1e-5, atol=1e-5)
    H.assert_close(st.returns, gold["ppo_returns"], "returns", rtol=1e-5, atol=1e-5)
    H.assert_close(st.advantages, gold["ppo_advantages"], "advantages", rtol=1e-4, atol=1e-4)
    H.assert_equal(st.dones, r["dones"].byte(), "dones")
    H.assert_equal(st.observations, r["obs"], "stored observations")
    vl, sl, ent, rnd, sym = alg.update(indices=torch.as_tensor(gold["ppo_perm"]).to(cuda))
    assert rnd is None and sym is None
    H.assert_close(torch.tensor([vl, sl, ent]), gold["ppo_losses"], "mean losses", rtol=1e-4, atol=1e-5)
    H.assert_close(torch.tensor([alg.learning_rate]), gold["ppo_lr_sequence"][-1:], "final learning rate", rtol=1e-6, atol=0)
    after = torch.cat([p.detach().flatten() for p in ac.parameters()])
    H.assert_close(after, gold["ppo_final_params"], "parameters after update", rtol=1e-4, atol=2e-5)
    assert alg.storage.step == 0
    # optimizer facade keeps torch.optim.Adam's checkpoint format
    sd = alg.optimizer.state_dict()
    assert set(sd) == {"state", "param_groups"} and float(sd["state"][0]["step"]) == 4.0
    assert sd["state"][1]["exp_avg"].shape == ac.actor[0].weight.shape
    alg.optimizer.load_state_dict(sd)


def test_state_dict_keys_match_reference_checkpoints(cuda, lt_lib):
    from locotouch_b200.loco_rl import ActorCritic

    ac = ActorCritic(270, 270, 12, [512, 256, 128], [512, 256, 128], "elu", 1.0)
    keys = list(ac.state_dict())
    assert keys[0] == "std"
    assert [k for k in keys if k.startswith("actor")] == [f"actor.{i}.{w}" for i in (0, 2, 4, 6) for w in ("weight", "bias")]
    assert [k for k in keys if k.startswith("critic")] == [f"critic.{i}.{w}" for i in (0, 2, 4, 6) for w in ("weight", "bias")]
    assert sum(p.numel() for p in ac.parameters()) == 607641  # locomotion actor-critic
    ac.to(cuda)
    flat, grads = ac.flatten_parameters()
    assert all(p.data_ptr() % 16 == 0 for p in ac.parameters())
    sd = {k: v.clone() for k, v in ac.state_dict().items()}
    ac.load_state_dict(sd)  # loading keeps the flat views
    assert ac.actor[0].weight.data_ptr() == flat.data_ptr() + ac._slices["actor.0.weight"][0] * 4
    with pytest.raises(ValueError):
        ActorCritic(10, 10, 4, noise_std_type="bogus")


def test_ppo_rejects_cpu_and_unused_branches(lt_lib):
    from locotouch_b200 import _C
    from locotouch_b200.loco_rl import PPO, ActorCritic

    ac = ActorCritic(8, 8, 4, [8], [8])
    with pytest.raises(_C.LocoTouchLibraryError):
        PPO(ac, device="cpu")
