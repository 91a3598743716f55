"""Scenario runners shared by the oracle tests (CPU) and the parity tests (GPU): same seeds as tests/golden/make_golden.py."""
from __future__ import annotations

import torch

from oracle import ppo as OP
from oracle.mdp import MdpOracle
from tests import helpers as H

PPO_SMALL = dict(T=24, N=32, obs_dim=270, A=12, hidden=[64, 48, 32], seed=5)
PPO_CFG = dict(num_learning_epochs=2, num_mini_batches=2, clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01,
               learning_rate=1.0e-3, max_grad_norm=1.0, desired_kl=0.01)
GAMMA, LAM = 0.99, 0.95


def run_oracle_mdp(scenario: str, on_step=None):
    """Runs the CPU oracle through the scenario; ``on_step(step, env, out, policy_obs, critic_obs, u_obs, u_euler)``."""
    spec, env, steps = H.make_mdp_env(scenario)
    oracle = MdpOracle(env, spec)
    for step in range(steps):
        out = oracle.step(env, auto_reset=True)
        u_obs, u_euler = H.mdp_noise(scenario, step, env.num_envs, spec.obs_dim_per_step)
        pol, cri = oracle.observe(env, u_noise=u_obs, u_obj_euler=u_euler)
        if on_step is not None:
            on_step(step, env, out, pol, cri, u_obs, u_euler, oracle)
        H.advance_mdp_env(env, step)
    return oracle, env


def oracle_ppo_rollout(golden):
    """Rebuilds the C1 rollout of the golden fixture with the oracle: act -> bootstrap -> GAE.  Returns flat storage."""
    c = PPO_SMALL
    T, N, A = c["T"], c["N"], c["A"]
    shapes = OP.actor_critic_shapes(c["obs_dim"], c["obs_dim"], A, c["hidden"], c["hidden"])
    flat = torch.as_tensor(golden["ppo_init_params"]).clone()
    params = OP.unflatten(flat, shapes)
    aw, ab, cw, cb = OP._split(params)
    r = H.make_rollout(T=T, N=N, obs_dim=c["obs_dim"], A=A, seed=c["seed"])
    eps = torch.as_tensor(golden["ppo_eps"])
    st = dict(obs=r["obs"], critic_obs=r["critic_obs"], actions=torch.zeros(T, N, A), logp=torch.zeros(T, N, 1), mu=torch.zeros(T, N, A),
              sigma=torch.zeros(T, N, A), values=torch.zeros(T, N, 1), rewards=torch.zeros(T, N, 1), dones=r["dones"].byte())
    with torch.no_grad():
        for t in range(T):
            mu = OP.mlp_forward(r["obs"][t], aw, ab)
            a, logp = OP.act_sample(mu, params["std"], eps[t])
            st["actions"][t], st["logp"][t, :, 0], st["mu"][t], st["sigma"][t] = a, logp, mu, params["std"].expand_as(mu)
            st["values"][t] = OP.mlp_forward(r["critic_obs"][t], cw, cb)
            st["rewards"][t, :, 0] = OP.bootstrap_rewards(r["rewards"][t, :, 0], st["values"][t], r["time_outs"][t, :, 0], GAMMA)
        last_values = OP.mlp_forward(r["critic_obs"][-1], cw, cb)
        st["returns"], st["advantages"] = OP.gae_returns(st["rewards"], st["values"], st["dones"], last_values, GAMMA, LAM, True)
    st["last_values"] = last_values
    st["time_outs"] = r["time_outs"]
    st["raw_rewards"] = r["rewards"]
    return flat, shapes, st
