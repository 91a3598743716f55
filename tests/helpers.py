"""Shared test drivers: seeded scenario generators used by the golden-vector script, the oracle tests and the GPU tests."""
from __future__ import annotations

import dataclasses
import os
from types import SimpleNamespace

import numpy as np
import torch

from locotouch_b200.mdp import task_spec as TS
from locotouch_b200.sim import synth
from locotouch_b200.sim.scene import SceneEntityCfg

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerance the north star states for floating point: <= 1e-5 relative (plus an absolute floor for values near zero)
RTOL, ATOL = 1e-5, 1e-6

MDP_SCENARIOS = {
    # name: (spec factory, num_envs, seed, steps, env kwargs)
    "locomotion": (TS.locomotion_spec, 101, 11, 36, dict()),
    "teacher": (TS.teacher_spec, 99, 12, 36, dict(with_object=True)),
}


def load_golden(name: str):
    path = os.path.join(GOLDEN_DIR, name)
    return dict(np.load(path, allow_pickle=False))


def make_mdp_env(scenario: str):
    spec_fn, n, seed, steps, kw = MDP_SCENARIOS[scenario]
    env = synth.make_env(n, seed=seed, **kw)
    return spec_fn(), env, steps


def mdp_noise(scenario: str, step: int, n: int, dps: int):
    """Explicit uniforms for the policy-group observation noise of one step (same draw for oracle, CUDA and reference)."""
    g = torch.Generator().manual_seed(7919 * (step + 1) + len(scenario))
    return torch.rand(n, dps, generator=g), torch.rand(n, 3, generator=g)


def advance_mdp_env(env, step: int):
    # keep the command for long stretches so that the gait state machine builds up valid air times, change it twice
    synth.advance(env, keep_cmd_prob=0.0 if step in (17, 29) else 1.0)


def gait_cfg(spec) -> SimpleNamespace:
    """RewardTermCfg-like object for the reference gait classes."""
    gp = {f.name: getattr(spec.gait, f.name) for f in dataclasses.fields(spec.gait) if f.name not in ("with_object", "obj_x_max", "obj_y_max")}
    return SimpleNamespace(params=dict(asset_cfg=SceneEntityCfg("robot"), sensor_cfg=SceneEntityCfg("robot_contact_senosr"), **gp))


def assert_close(actual, expected, what="", rtol=RTOL, atol=ATOL):
    a = torch.as_tensor(np.asarray(actual.detach().cpu() if torch.is_tensor(actual) else actual)).double()
    e = torch.as_tensor(np.asarray(expected.detach().cpu() if torch.is_tensor(expected) else expected)).double()
    assert a.shape == e.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    err = (a - e).abs()
    tol = atol + rtol * e.abs()
    bad = err > tol
    if bad.any():
        i = int(torch.argmax(err - tol))
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.numel()} elements exceed rtol={rtol} atol={atol}; worst |{a.flatten()[i]:.9g} - {e.flatten()[i]:.9g}| = {err.flatten()[i]:.3g}"
        )


def assert_equal(actual, expected, what=""):
    a = torch.as_tensor(np.asarray(actual.detach().cpu() if torch.is_tensor(actual) else actual))
    e = torch.as_tensor(np.asarray(expected.detach().cpu() if torch.is_tensor(expected) else expected))
    assert a.shape == e.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    neq = a.to(torch.int64) != e.to(torch.int64) if a.dtype != e.dtype else a != e
    assert not bool(neq.any()), f"{what}: {int(neq.sum())}/{neq.numel()} elements differ (bit-exact comparison)"


ACTION_TERM = dict(n=37, J=12, steps=10, seed=23, clip=100.0, raw_scale=0.25, scale=0.5)


def action_term_tape(c=ACTION_TERM):
    """Seeded inputs of the action-term fixture: per-step policy actions (one step far outside the clip range), the per-env joint
    offsets and the env ids reset after each step (none on some steps, a single env, a random subset, all)."""
    g = torch.Generator().manual_seed(c["seed"])
    n, J, steps = c["n"], c["J"], c["steps"]
    actions = torch.randn(steps, n, J, generator=g)
    actions[3] *= 300.0
    offset = torch.randn(n, J, generator=g)
    resets = []
    for s in range(steps):
        if s % 4 == 0:
            ids = torch.zeros(0, dtype=torch.long)
        elif s == 5:
            ids = torch.tensor([n - 1])
        elif s == 7:
            ids = torch.arange(n)
        else:
            ids = (torch.rand(n, generator=g) < 0.3).nonzero().flatten()
        resets.append(ids)
    return actions, offset, resets


# ----------------------------------------------------------------------------------------------------- C1: loco_rl PPO
def make_rollout(T=24, N=64, obs_dim=270, A=12, seed=0):
    """Synthetic rollout tensors of config C1."""
    g = torch.Generator().manual_seed(seed)
    r = dict(
        obs=torch.randn(T, N, obs_dim, generator=g),
        critic_obs=torch.randn(T, N, obs_dim, generator=g),
        rewards=torch.randn(T, N, 1, generator=g) * 0.02,
        values=torch.randn(T, N, 1, generator=g) * 0.5,
        last_values=torch.randn(N, 1, generator=g) * 0.5,
        dones=(torch.rand(T, N, 1, generator=g) < 0.02),
    )
    r["time_outs"] = r["dones"] & (torch.rand(T, N, 1, generator=g) < 0.5)
    # edge cases: an env that is done at every step, one done at the last step only
    r["dones"][:, 0] = True
    r["dones"][:, 1] = False
    r["dones"][-1, 1] = True
    return r


class TapeEnv:
    """Deterministic stand-in for the IsaacLab env of ReplayBuffer tests: replays pre-generated observations, rewards and
    dones whatever the actions are (the reference's ReplayBuffer and the drop-in consume the same tape).  ``rsl_style``:
    the tuple API ``evaluate`` uses (get_observations() -> (obs, extras); step -> (obs, reward, dones, extras))."""

    def __init__(self, num_envs=24, steps=200, device="cpu", seed=0, p_done=0.04, obs_dim=30, tactile_dim=16, rsl_style=False):
        g = torch.Generator().manual_seed(seed)
        self.num_envs, self.device, self.rsl_style = num_envs, device, rsl_style
        self.policy = torch.randn(steps + 1, num_envs, obs_dim, generator=g).to(device)
        self.tactile = (torch.rand(steps + 1, num_envs, tactile_dim, generator=g) < 0.2).float().to(device)
        self.rewards = (torch.randn(steps, num_envs, generator=g) * 0.1).to(device)
        dones = torch.rand(steps, num_envs, generator=g) < p_done
        dones[:, 0] = True    # an env that finishes every step (length-1 trajectories)
        if num_envs > 1:
            dones[:, 1] = False   # one that never finishes (never recorded)
        self.dones = dones.to(device)
        self.t = 0
        self.resets = 0

    def _obs(self):
        if self.rsl_style:
            return self.policy[self.t], {"observations": {"tactile": self.tactile[self.t]}}
        return {"policy": self.policy[self.t], "tactile": self.tactile[self.t]}

    def get_observations(self):
        return self._obs()

    def reset(self):
        self.resets += 1

    def step(self, action):
        r, d = self.rewards[self.t], self.dones[self.t]
        self.t += 1
        if self.rsl_style:
            obs, extras = self._obs()
            return obs, r, d, extras
        return self._obs(), r, d, {}
