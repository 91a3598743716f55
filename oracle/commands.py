"""ORACLE (test infrastructure, never imported by the product path).

CPU (torch fp32 / Python double) restatement of the LocoTouch velocity-command term and velocity curriculum
(SURVEY.md 8f rank 3):

  CommandOracle        UniformVelocityCommandGaitLoggingMultiSampling   reference locotouch/mdp/commands.py:379-576
                       over [IL] CommandTerm / UniformVelocityCommand   (oracle/il_commands.py, restated, UNPINNED)
  VelCurriculumOracle  ModifyVelCommandsRangeBasedonReward              reference locotouch/mdp/curriculums.py:184-274

Randomness goes through a small ``rng`` object so that the SAME statements serve two purposes:
  * ``TorchRng``    issues exactly the torch calls of the reference in the reference's order (``uniform_`` on a fresh
                    ``torch.empty(n)``, ``torch.multinomial``, ``torch.randint``): under one ``torch.manual_seed`` the oracle
                    and the unmodified reference classes produce identical tensors -- this is how the restatement is pinned
                    (tests/golden/commands_c3.npz, tests/test_commands.py).
  * ``ExplicitRng`` takes per-env uniforms ``u[N, 8]`` (slot 0 time_left, 1-3 x / y / yaw value, 4-6 x / y / yaw bin,
                    7 standing) -- the parity mode of the CUDA kernel (lt_command_step), which cannot share torch's CPU
                    stream.  Value arithmetic is torch's: ``fma(u, hi - lo, lo)`` in fp32 with fp32-rounded bounds
                    (checked against ``Tensor.uniform_`` in tests/test_commands.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

SLOT_TIME, SLOT_X, SLOT_Y, SLOT_Z, SLOT_BIN_X, SLOT_BIN_Y, SLOT_BIN_Z, SLOT_STAND = range(8)


class TorchRng:
    def uniform(self, env_ids, lo, hi, slot):
        return torch.empty(len(env_ids)).uniform_(lo, hi)

    def multinomial(self, probs, env_ids, slot):
        return torch.multinomial(probs, len(env_ids), replacement=True)

    def randint(self, high, env_ids, slot):
        return torch.randint(0, high, (len(env_ids),))


class ExplicitRng:
    def __init__(self, u: torch.Tensor | None = None):
        self.u = u

    def uniform(self, env_ids, lo, hi, slot):
        # torch's CPU uniform_ transform is ONE fused multiply-add in fp32: fma(u, hi32 - lo32, lo32).  Emulated through double
        # (the product of two 24-bit significands is exact there)
        lo32, hi32 = torch.tensor(lo, dtype=torch.float32), torch.tensor(hi, dtype=torch.float32)
        return (self.u[env_ids, slot].double() * (hi32 - lo32).double() + lo32.double()).float()

    def multinomial(self, probs, env_ids, slot):
        c = torch.cumsum(probs, 0) / probs.sum()  # fp32, like the normalised distribution torch.multinomial searches
        u = self.u[env_ids, slot]
        return (u >= c[0]).long() + (u >= c[1]).long()

    def randint(self, high, env_ids, slot):
        return torch.clamp((self.u[env_ids, slot] * high).long(), max=high - 1)


def _bins(cur, prev):
    """commands.py:439-447 / 475-478: (new low, old low), (old low, old high), (old high, new high) as an fp32 tensor."""
    return torch.tensor([(cur[0], prev[0]), (prev[0], prev[1]), (prev[1], cur[1])], dtype=torch.float32)


class CommandOracle:
    """State and methods named like the reference class; ``cfg`` is a plain dict of its cfg fields."""

    def __init__(self, env, *, ranges, resampling_time_range=(8.0, 8.0), rel_standing_envs=0.1, final_rel_standing_envs=0.0,
                 new_command_probs=0.15, initial_zero_command_steps=0, final_initial_zero_command_steps=0,
                 binary_maximal_command=False, feet_sensor_ids=(13, 14, 15, 16), gait_valid_last_air_time=None, rng=None,
                 cast_maximal_to_fp32=False):
        n = env.num_envs
        self.env, self.rng = env, rng or TorchRng()
        self.ranges = {k: tuple(v) for k, v in ranges.items()}  # lin_vel_x / lin_vel_y / ang_vel_z
        self.previous_ranges = {k: tuple(v) for k, v in ranges.items()}  # commands.py:431-433
        self.equal = {k: True for k in ranges}  # :434-436
        self.resampling_time_range = resampling_time_range
        self.rel_standing_envs, self.final_rel_standing_envs = rel_standing_envs, final_rel_standing_envs
        self.sampling_probs = torch.tensor([new_command_probs, 1.0 - 2 * new_command_probs, new_command_probs])  # :448
        self.sampling_ranges = {k: _bins(self.ranges[k], self.previous_ranges[k]) for k in ranges}
        self.initial_zero_command_steps = initial_zero_command_steps  # :451
        self.final_initial_zero_command_steps = final_initial_zero_command_steps
        self.binary_maximal_command = binary_maximal_command
        # reference quirk: after the curriculum has run, cfg.ranges hold np.float64 (np.clip) and commands.py:520-521 builds a float64
        # tensor that index_put_ rejects.  False reproduces that (RuntimeError); True rounds to fp32, which is what the CUDA path does.
        self.cast_maximal_to_fp32 = cast_maximal_to_fp32
        self.maximal_command_sampling = torch.tensor([[i, j, k] for i in (-1, 1) for j in (-1, 1) for k in (-1, 1)], dtype=torch.float32)  # :453-463
        self.feet = list(feet_sensor_ids)
        self.gait_vla = gait_valid_last_air_time  # callable -> [N, 4] tensor (rewards.py:99), or None
        # [IL] CommandTerm / UniformVelocityCommand state
        self.vel_command_b = torch.zeros(n, 3)
        self.vel_command_b_buffer = torch.zeros(n, 3)  # :450
        self.time_left = torch.zeros(n)
        self.command_counter = torch.zeros(n, dtype=torch.long)
        self.is_standing_env = torch.zeros(n, dtype=torch.bool)
        z = lambda: torch.zeros(n)  # noqa: E731
        self.metrics = {"error_vel_xy": z(), "error_vel_yaw": z()}  # [IL]
        for k in ("foot_air_time_variance", "foot_step_frequency", "pair_1_step_frequency", "pair_2_step_frequency", "step_air_time",
                  "pair_1_air_time", "pair_2_air_time"):  # :385-391
            self.metrics[k] = z()
        for k in ("lin_vel_x", "lin_vel_y", "ang_vel_z"):  # :465-467
            self.metrics[k] = z()
        self.metrics["initial_zero_command_steps"] = torch.ones(n) * initial_zero_command_steps  # :468
        self.metrics["rel_standing_envs"] = torch.ones(n) * rel_standing_envs  # :469

    @property
    def command(self):
        return self.vel_command_b

    # ------------------------------------------------------------------ commands.py:471-505
    def set_ranges(self, lin_vel_x=None, lin_vel_y=None, ang_vel_z=None):
        for key, new in (("lin_vel_x", lin_vel_x), ("lin_vel_y", lin_vel_y), ("ang_vel_z", ang_vel_z)):
            if new is not None:
                self.previous_ranges[key] = tuple(self.ranges[key])
                self.ranges[key] = tuple(new)
                self.equal[key] = self.previous_ranges[key] == self.ranges[key]
                self.sampling_ranges[key] = _bins(self.ranges[key], self.previous_ranges[key])
        if all(self.equal.values()):  # :497-502
            self.initial_zero_command_steps = self.final_initial_zero_command_steps
            self.rel_standing_envs = self.final_rel_standing_envs

    # ------------------------------------------------------------------ commands.py:393-418 + 507-515
    def _update_metrics(self):
        env, m = self.env, self.metrics
        rb = env.scene["robot"].data
        m["error_vel_xy"] = torch.linalg.norm(self.vel_command_b[:, :2] - rb.root_lin_vel_b[:, :2], dim=-1)
        m["error_vel_yaw"] = torch.abs(self.vel_command_b[:, 2] - rb.root_ang_vel_b[:, 2])
        last_air = env.scene.sensors["robot_contact_senosr"].data.last_air_time[:, self.feet]
        m["foot_air_time_variance"] = torch.var(last_air, dim=1)
        if self.gait_vla is not None:
            vla = self.gait_vla()
            masked = vla[torch.all(vla > 1.0e-6, dim=1)]
            avg = torch.mean(masked)
            m["foot_step_frequency"][:] = (1.0 / avg / 2.0) if avg > 0 else 0.0
            p1, p2 = torch.mean(masked[:, [0, 1]]), torch.mean(masked[:, [2, 3]])
            m["pair_1_step_frequency"][:] = (1.0 / p1 / 2.0) if p1 > 0 else 0.0
            m["pair_2_step_frequency"][:] = (1.0 / p2 / 2.0) if p2 > 0 else 0.0
            m["step_air_time"][:] = avg if avg > 0 else 0.0
            m["pair_1_air_time"][:] = p1 if p1 > 0 else 0.0
            m["pair_2_air_time"][:] = p2 if p2 > 0 else 0.0
        m["lin_vel_x"][:] = self.ranges["lin_vel_x"][1]
        m["lin_vel_y"][:] = self.ranges["lin_vel_y"][1]
        m["ang_vel_z"][:] = self.ranges["ang_vel_z"][1]
        m["initial_zero_command_steps"][:] = self.initial_zero_command_steps
        m["rel_standing_envs"][:] = self.rel_standing_envs

    # ------------------------------------------------------------------ commands.py:517-559
    def _resample_command(self, env_ids):
        rng = self.rng
        if self.binary_maximal_command:
            idx = rng.randint(self.maximal_command_sampling.shape[0], env_ids, SLOT_X)
            maximal = torch.tensor([self.ranges["lin_vel_x"][1], self.ranges["lin_vel_y"][1], self.ranges["ang_vel_z"][1]])
            if self.cast_maximal_to_fp32:
                maximal = maximal.float()
            self.vel_command_b[env_ids] = self.maximal_command_sampling[idx] * maximal
        else:
            dims = (("lin_vel_x", 0, SLOT_X, SLOT_BIN_X), ("lin_vel_y", 1, SLOT_Y, SLOT_BIN_Y), ("ang_vel_z", 2, SLOT_Z, SLOT_BIN_Z))
            if all(self.equal.values()):  # [IL] UniformVelocityCommand._resample_command
                for key, d, slot, _ in dims:
                    self.vel_command_b[env_ids, d] = rng.uniform(env_ids, *self.ranges[key], slot)
            else:
                for key, d, slot, bslot in dims:
                    if self.equal[key]:
                        self.vel_command_b[env_ids, d] = rng.uniform(env_ids, *self.ranges[key], slot)
                    else:
                        bin_indices = rng.multinomial(self.sampling_probs, env_ids, bslot)
                        for i in range(3):
                            sel = bin_indices == i
                            if sel.any():
                                chosen = env_ids[torch.nonzero(sel, as_tuple=True)[0]]
                                low, high = float(self.sampling_ranges[key][i, 0]), float(self.sampling_ranges[key][i, 1])
                                self.vel_command_b[chosen, d] = rng.uniform(chosen, low, high, slot)
            self.is_standing_env[env_ids] = rng.uniform(env_ids, 0.0, 1.0, SLOT_STAND) <= self.rel_standing_envs
        self.vel_command_b_buffer[env_ids] = self.vel_command_b[env_ids].clone()  # :558
        self._set_zero_command_for_beginning_steps()  # :559

    def _set_zero_command_for_beginning_steps(self):  # :566-570
        ids = (self.env.episode_length_buf < self.initial_zero_command_steps).nonzero(as_tuple=True)[0]
        if len(ids):
            self.vel_command_b[ids] = self.vel_command_b_buffer[ids] * 0.0

    def _recover_command_for_beginning_steps(self):  # :572-576
        ids = (self.env.episode_length_buf == self.initial_zero_command_steps).nonzero(as_tuple=True)[0]
        if len(ids):
            self.vel_command_b[ids] = self.vel_command_b_buffer[ids].clone()

    def _update_command(self):  # :561-564 + [IL] standing envs
        self._set_zero_command_for_beginning_steps()
        self._recover_command_for_beginning_steps()
        self.vel_command_b[self.is_standing_env.nonzero(as_tuple=False).flatten(), :] = 0.0

    # ------------------------------------------------------------------ [IL] CommandTerm drivers
    def _resample(self, env_ids):
        if len(env_ids) != 0:
            self.time_left[env_ids] = self.rng.uniform(env_ids, *self.resampling_time_range, SLOT_TIME)
            self._resample_command(env_ids)
            self.command_counter[env_ids] += 1

    def reset(self, env_ids=None):
        env_ids = torch.arange(self.env.num_envs) if env_ids is None else torch.as_tensor(env_ids, dtype=torch.long)
        extras = {}
        for name, value in self.metrics.items():
            extras[name] = torch.mean(value[env_ids]).item()
            value[env_ids] = 0.0
        self.command_counter[env_ids] = 0
        self._resample(env_ids)
        return extras

    def compute(self, dt: float):
        self._update_metrics()
        self.time_left -= dt
        self._resample((self.time_left <= 0.0).nonzero().flatten())
        self._update_command()


class VelCurriculumOracle:
    """ModifyVelCommandsRangeBasedonReward (curriculums.py:184-274).  ``episode_sums`` maps reward name -> [N] tensor
    ([IL] RewardManager._episode_sums); ``weights`` / ``sigmas`` are the two tracking terms' cfg values."""

    def __init__(self, env, command: CommandOracle, episode_sums, *, weight_lin, sigma_lin, weight_ang, sigma_ang,
                 command_maximum_ranges=(0.6, 0.3, math.pi / 4), curriculum_bins=(20, 20, 20), reset_envs_episode_length=0.95,
                 reward_name_lin="track_lin_vel_xy", reward_name_ang="track_ang_vel_z", error_threshold_lin=0.05, error_threshold_ang=0.08,
                 repeat_times_lin=5, repeat_times_ang=5, max_distance_bins=3):
        n = env.num_envs
        self.env, self.cmd, self.sums = env, command, episode_sums
        self.max = list(command_maximum_ranges)
        r = command.ranges
        self.exp = [(self.max[0] - r["lin_vel_x"][1]) / curriculum_bins[0], (self.max[1] - r["lin_vel_y"][1]) / curriculum_bins[1],
                    (self.max[2] - r["ang_vel_z"][1]) / curriculum_bins[2]]  # :191-193
        self.reset_len = reset_envs_episode_length * env.max_episode_length_s  # :194 (sic: seconds against a step count)
        self.name_lin, self.name_ang = reward_name_lin, reward_name_ang
        self.thr_lin = math.exp(-error_threshold_lin / sigma_lin) * weight_lin * env.max_episode_length_s  # :199
        self.thr_ang = math.exp(-error_threshold_ang / sigma_ang) * weight_ang * env.max_episode_length_s  # :200
        self.rep_lin, self.rep_ang, self.max_dist = repeat_times_lin, repeat_times_ang, max_distance_bins
        self.lin_forward_bins = self.ang_forward_bins = 0
        self.success_lin = self.success_ang = 0
        self.reseted_lin, self.reseted_ang = torch.zeros(n, dtype=torch.bool), torch.zeros(n, dtype=torch.bool)
        self.len_lin, self.len_ang, self.sum_lin, self.sum_ang = torch.zeros(n), torch.zeros(n), torch.zeros(n), torch.zeros(n)

    def __call__(self, env_ids):
        env, c, r = self.env, self.cmd, self.cmd.ranges
        env_ids = torch.as_tensor(env_ids, dtype=torch.long)
        if (r["lin_vel_x"][1] != self.max[0] or not c.equal["lin_vel_x"] or r["lin_vel_y"][1] != self.max[1] or not c.equal["lin_vel_y"]) \
                and self.lin_forward_bins - self.ang_forward_bins <= self.max_dist:  # :229-231
            self.reseted_lin[env_ids] = True
            self.len_lin[env_ids] = env.episode_length_buf[env_ids].float()
            self.sum_lin[env_ids] = self.sums[self.name_lin][env_ids]
            if torch.all(self.reseted_lin) and torch.mean(self.len_lin) > self.reset_len and torch.mean(self.sum_lin) > self.thr_lin:
                self.success_lin += 1
                if self.success_lin == self.rep_lin:
                    lx = np.clip(r["lin_vel_x"][0] - self.exp[0], -self.max[0], 0.0)
                    ly = np.clip(r["lin_vel_y"][0] - self.exp[1], -self.max[1], 0.0)
                    c.set_ranges(lin_vel_x=(lx, -lx), lin_vel_y=(ly, -ly), ang_vel_z=None)
                    self.success_lin = 0
                    self.lin_forward_bins += 1
                self.reseted_lin[:] = False
                self.len_lin[:] = 0
                self.sum_lin[:] = 0
        r = self.cmd.ranges
        if (r["ang_vel_z"][1] != self.max[2] or not c.equal["ang_vel_z"]) and self.ang_forward_bins - self.lin_forward_bins <= self.max_dist:  # :252-253
            self.reseted_ang[env_ids] = True
            self.len_ang[env_ids] = env.episode_length_buf[env_ids].float()
            self.sum_ang[env_ids] = self.sums[self.name_ang][env_ids]
            if torch.all(self.reseted_ang) and torch.mean(self.len_ang) > self.reset_len and torch.mean(self.sum_ang) > self.thr_ang:
                self.success_ang += 1
                if self.success_ang == self.rep_ang:
                    lz = np.clip(r["ang_vel_z"][0] - self.exp[2], -self.max[2], 0.0)
                    c.set_ranges(lin_vel_x=None, lin_vel_y=None, ang_vel_z=(lz, -lz))
                    self.success_ang = 0
                    self.ang_forward_bins += 1
                self.reseted_ang[:] = False
                self.len_ang[:] = 0
                self.sum_ang[:] = 0
