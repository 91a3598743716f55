"""Recurrent encoder policies + the recurrent branch of PPO.update on the GPU against one full iteration of
the UNMODIFIED reference (golden recurrent_ppo_c1.npz: ActorCriticRNNEncoder and ActorCriticPreEncoderRNNEncoder; rollout with
the hidden-state bookkeeping -> compute_returns -> 2 epochs x 2 recurrent mini-batches with the adaptive learning rate)."""
import pytest
import torch

from tests import helpers as H
from tests.golden.make_golden import RECURRENT_PPO, recurrent_policy_kwargs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["rnn", "pre_rnn"])
def test_recurrent_policy_iteration_matches_reference(cuda, lt_lib, kind):
    from locotouch_b200.loco_rl import PPO, ActorCriticPreEncoderRNNEncoder, ActorCriticRNNEncoder

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    gold = H.load_golden("recurrent_ppo_c1.npz")
    c = RECURRENT_PPO
    T, N, A = c["T"], c["N"], c["A"]
    obs_dim = c["flat_dim"] + c["enc_dim"]
    cls = ActorCriticRNNEncoder if kind == "rnn" else ActorCriticPreEncoderRNNEncoder
    ac = cls(**recurrent_policy_kwargs(kind))
    assert ac.is_recurrent
    assert [k for k, _ in ac.named_parameters()] == [str(n) for n in gold[f"{kind}_names"]], "state_dict keys must match reference checkpoints"
    flat, off = torch.as_tensor(gold[f"{kind}_init"]), 0
    with torch.no_grad():
        for p in ac.parameters():
            p.copy_(flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
    alg = PPO(ac, num_learning_epochs=c["num_learning_epochs"], num_mini_batches=c["num_mini_batches"], clip_param=0.2, gamma=0.99, lam=0.95,
              value_loss_coef=1.0, entropy_coef=0.01, learning_rate=1.0e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
              desired_kl=0.01, device="cuda:0")
    alg.init_storage(N, T, [obs_dim], [obs_dim], [A])
    r = H.make_rollout(T=T, N=N, obs_dim=obs_dim, A=A, seed=c["seed"])
    eps = torch.as_tensor(gold[f"{kind}_eps"]).to(cuda)
    for t in range(T):
        ac.rng = lambda mean, _e=eps[t]: _e
        alg.act(r["obs"][t].to(cuda), r["critic_obs"][t].to(cuda))
        alg.process_env_step(r["rewards"][t, :, 0].to(cuda), r["dones"][t, :, 0].long().to(cuda), {"time_outs": r["time_outs"][t, :, 0].to(cuda)})
    alg.compute_returns(r["critic_obs"][-1].to(cuda))
    st = alg.storage
    H.assert_close(st.actions, gold[f"{kind}_actions"], "actions", rtol=1e-5, atol=2e-5)
    H.assert_close(st.actions_log_prob, gold[f"{kind}_logp"], "log prob", rtol=1e-5, atol=1e-4)
    H.assert_close(st.values, gold[f"{kind}_values"], "values", rtol=1e-5, atol=2e-5)
    H.assert_close(st.returns, gold[f"{kind}_returns"], "returns", rtol=1e-5, atol=2e-5)
    H.assert_close(st.advantages, gold[f"{kind}_advantages"], "advantages", rtol=1e-4, atol=2e-4)
    H.assert_close(st.saved_hidden_states_a[0], gold[f"{kind}_hid_a"], "saved actor GRU states", rtol=1e-5, atol=1e-5)
    H.assert_close(st.saved_hidden_states_c[0], gold[f"{kind}_hid_c"], "saved critic GRU states", rtol=1e-5, atol=1e-5)
    H.assert_close(ac.get_hidden_states()[0], gold[f"{kind}_final_hidden"], "actor GRU state after the rollout (done envs cleared)", rtol=1e-5, atol=1e-5)
    vl, sl, ent, rnd, sym = alg.update()
    assert rnd is None and sym is None
    H.assert_close(torch.tensor([vl, sl, ent]), gold[f"{kind}_losses"], "mean losses", rtol=2e-4, atol=2e-5)
    H.assert_close(torch.tensor([alg.learning_rate]), gold[f"{kind}_lr_sequence"][-1:], "final learning rate (adaptive schedule)", rtol=1e-6, atol=0)
    after = torch.cat([p.detach().flatten() for p in ac.parameters()])
    H.assert_close(after, gold[f"{kind}_final"], "parameters after the recurrent update", rtol=2e-4, atol=3e-5)
    assert float((after.cpu() - flat).abs().max()) > 1e-4 and alg.storage.step == 0


def test_unpad_is_differentiable(cuda, lt_lib):
    from locotouch_b200.loco_rl.utils import split_and_pad_trajectories, unpad_trajectories

    g = torch.Generator().manual_seed(3)
    x = torch.randn(24, 17, 5, generator=g).to(cuda)
    dones = (torch.rand(24, 17, 1, generator=g) < 0.1).to(cuda)
    padded, masks = split_and_pad_trajectories(x, dones)
    p = padded.clone().requires_grad_(True)
    w = torch.randn(24, 17, 5, generator=g).to(cuda)
    (unpad_trajectories(p, masks) * w).sum().backward()
    want, _ = split_and_pad_trajectories(w, dones)
    H.assert_equal(p.grad, want, "gradient of unpad = zero-padded scatter of the incoming gradient")
