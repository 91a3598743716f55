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


# ------------------------------------------------------------------------- commands + velocity curriculum
import math  # noqa: E402
from types import SimpleNamespace  # noqa: E402

from locotouch_b200.sim import synth  # noqa: E402

CMD_SCENARIO = dict(N=96, steps=70, seed=21, dt=0.02)
CMD_CFG = dict(ranges=dict(lin_vel_x=(-0.2, 0.2), lin_vel_y=(-0.1, 0.1), ang_vel_z=(-math.pi / 10, math.pi / 10)),
               resampling_time_range=(0.04, 0.2),  # a few steps, so that timer-driven resampling happens inside the scenario
               rel_standing_envs=0.1, final_rel_standing_envs=0.05, new_command_probs=0.15, initial_zero_command_steps=0,
               final_initial_zero_command_steps=3)
CUR_CFG = dict(command_maximum_ranges=[0.5, 0.25, math.pi / 4], curriculum_bins=[2, 2, 2], reset_envs_episode_length=0.98,
               error_threshold_lin=0.08, error_threshold_ang=0.1, repeat_times_lin=2, repeat_times_ang=1, max_distance_bins=1)
CMD_REWARD_CFG = dict(track_lin_vel_xy=dict(weight=1.0, sigma=0.25), track_ang_vel_z=dict(weight=0.5, sigma=0.25))
FEET_SENSOR_IDS = [13, 14, 15, 16]


class CommandScenario:
    """Deterministic tape around a command term + curriculum term: mutates the env state the two terms read and decides which
    envs reset at every step, in the order of [IL] ManagerBasedRLEnv.step (episode length += 1; curriculum -> command reset for
    the reset envs -> their episode length = 0; command compute).  Its own generator: the global torch RNG is left to the terms."""

    def __init__(self, binary_maximal_command=False, device="cpu"):
        c = CMD_SCENARIO
        self.N, self.steps, self.dt = c["N"], c["steps"], c["dt"]
        self.gen = torch.Generator().manual_seed(c["seed"])
        self.device = torch.device(device)
        self.env = synth.make_env(self.N, seed=c["seed"])
        if self.device.type != "cpu":
            self.env = self.env.to(self.device)  # same bits; the tape below is generated on the CPU and copied
        self.env.episode_length_buf[:] = 0
        self.vla = torch.zeros(self.N, 4, device=self.device)  # gait term state read by the logging metrics (rewards.py:99)
        self.sums = {k: torch.zeros(self.N, device=self.device) for k in CMD_REWARD_CFG}
        self.env.reward_manager._episode_sums = self.sums
        for k, v in CMD_REWARD_CFG.items():
            self.env.reward_manager.set_term_cfg(k, SimpleNamespace(weight=v["weight"], params={"sigma": v["sigma"]}))
        self.env.reward_manager.set_term_cfg("gait", SimpleNamespace(func=SimpleNamespace(valid_last_air_time=self.vla)))
        self.binary = binary_maximal_command
        s = self.env.max_episode_length_s
        self.thr = {k: math.exp(-CUR_CFG["error_threshold_" + k.split("_")[1]] / v["sigma"]) * v["weight"] * s for k, v in CMD_REWARD_CFG.items()}

    def uniforms(self, t: int, phase: int):
        """Explicit per-env uniforms [N, 8] of call ``phase`` (0 reset, 1 compute) at step t (t = -1: the initial reset)."""
        g = torch.Generator().manual_seed(104729 * (t + 2) + phase)
        return torch.rand(self.N, 8, generator=g)

    def mutate(self, t: int):
        """New robot velocities / contact timers / gait state / episode sums; returns the env ids that reset at this step."""
        env, g, N = self.env, self.gen, self.N
        env.episode_length_buf += 1
        rb = env.scene["robot"].data
        rb.root_lin_vel_b = (torch.randn(N, 3, generator=g) * 0.5).to(self.device)
        rb.root_ang_vel_b = (torch.randn(N, 3, generator=g) * 0.5).to(self.device)
        sensor = env.scene.sensors["robot_contact_senosr"].data
        sensor.last_air_time = (torch.rand(N, sensor.last_air_time.shape[1], generator=g) * 0.6).to(self.device)
        # gait state: nothing valid on the first steps (empty mask -> nan means), then a growing share of envs with four valid feet
        vla = torch.rand(N, 4, generator=g) * 0.4
        vla[torch.rand(N, 4, generator=g) < (1.0 if t < 2 else 0.15)] = 0.0
        self.vla.copy_(vla)
        level = 0.85 if (t // 6) % 3 == 2 else 1.15  # every third window the tracking rewards fall short of the threshold
        for k in self.sums:
            self.sums[k].copy_((0.8 + 0.4 * torch.rand(N, generator=g)) * level * self.thr[k])
        ids = (torch.rand(N, generator=g) < 0.3).nonzero().flatten()
        if t % 5 == 4:  # every env resets at least once per window: torch.all(env_reseted) becomes true
            ids = torch.arange(N)
        if t == 7:
            ids = ids[:0]  # a step without resets
        return ids.to(self.device)

    def run(self, command, curriculum, snapshot, explicit_rng=None):
        """``command.reset / compute`` and ``curriculum(env, env_ids)`` are duck-typed (reference classes, oracle, CUDA drop-in).
        ``snapshot(t, extras)`` records.  ``explicit_rng``: object whose ``.u`` is set before each call (parity mode)."""
        env = self.env
        if explicit_rng is not None:
            explicit_rng.u = self.uniforms(-1, 0)
        extras = command.reset(torch.arange(self.N, device=self.device))
        snapshot(-1, extras)
        for t in range(self.steps):
            ids = self.mutate(t)
            extras = None
            if len(ids):
                env.episode_length_buf[ids] += 25  # the tape's stand-in for long episodes (curriculum needs mean length > 19.6)
                curriculum(env, ids)
                if explicit_rng is not None:
                    explicit_rng.u = self.uniforms(t, 0)
                extras = command.reset(ids)
                env.episode_length_buf[ids] = 0
            if explicit_rng is not None:
                explicit_rng.u = self.uniforms(t, 1)
            command.compute(self.dt)
            snapshot(t, extras)


CMD_METRICS = ("error_vel_xy", "error_vel_yaw", "foot_air_time_variance", "foot_step_frequency", "pair_1_step_frequency", "pair_2_step_frequency",
               "step_air_time", "pair_1_air_time", "pair_2_air_time", "lin_vel_x", "lin_vel_y", "ang_vel_z", "initial_zero_command_steps",
               "rel_standing_envs")


class CommandRecorder:
    """Collects per-step snapshots of a command term into stacked arrays; ``view(cmd)`` adapts the three implementations."""

    def __init__(self, view):
        self.view, self.rows = view, []

    def __call__(self, t, extras):
        v = self.view()
        cpu = lambda x: torch.as_tensor(x).detach().cpu().clone()  # noqa: E731
        row = dict(cmd=cpu(v["cmd"]), buffer=cpu(v["buffer"]), time_left=cpu(v["time_left"]), standing=cpu(v["standing"]),
                   counter=cpu(v["counter"]), metrics=torch.stack([cpu(v["metrics"][k]).float().expand(v["cmd"].shape[0]).clone() for k in CMD_METRICS]),
                   scalars=torch.tensor(list(v["ranges"]) + list(v["previous"]) + [float(x) for x in v["equal"]] + [float(v["izcs"]), float(v["rel_standing"])]
                                        + [float(x) for x in v["curriculum"]], dtype=torch.float64),
                   extras=torch.tensor([float("nan") if extras is None else float(extras[k]) for k in CMD_METRICS], dtype=torch.float64))
        self.rows.append(row)

    def stacked(self):
        return {k: torch.stack([r[k] for r in self.rows]) for k in self.rows[0]}


def oracle_command_view(cmd, cur):
    def view():
        r, p = cmd.ranges, cmd.previous_ranges
        keys = ("lin_vel_x", "lin_vel_y", "ang_vel_z")
        return dict(cmd=cmd.vel_command_b, buffer=cmd.vel_command_b_buffer, time_left=cmd.time_left, standing=cmd.is_standing_env,
                    counter=cmd.command_counter, metrics=cmd.metrics, ranges=[x for k in keys for x in r[k]], previous=[x for k in keys for x in p[k]],
                    equal=[cmd.equal[k] for k in keys], izcs=cmd.initial_zero_command_steps, rel_standing=cmd.rel_standing_envs,
                    curriculum=[cur.lin_forward_bins, cur.ang_forward_bins, cur.success_lin, cur.success_ang, int(cur.reseted_lin.sum()),
                                int(cur.reseted_ang.sum()), float(cur.len_lin.sum()), float(cur.len_ang.sum())])
    return view


def make_oracle_command(sc: CommandScenario, rng, cast_maximal_to_fp32=False):
    from oracle.commands import CommandOracle, VelCurriculumOracle

    cmd = CommandOracle(sc.env, **CMD_CFG, binary_maximal_command=sc.binary, feet_sensor_ids=FEET_SENSOR_IDS, gait_valid_last_air_time=lambda: sc.vla, rng=rng,
                        cast_maximal_to_fp32=cast_maximal_to_fp32)
    w = CMD_REWARD_CFG
    cur = VelCurriculumOracle(sc.env, cmd, sc.sums, weight_lin=w["track_lin_vel_xy"]["weight"], sigma_lin=w["track_lin_vel_xy"]["sigma"],
                              weight_ang=w["track_ang_vel_z"]["weight"], sigma_ang=w["track_ang_vel_z"]["sigma"], **CUR_CFG)
    return cmd, (lambda env, ids: cur(ids)), cur
