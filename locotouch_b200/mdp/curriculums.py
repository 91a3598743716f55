"""Velocity curriculum term with the reference's class name and manager-term signature
(reference locotouch/mdp/curriculums.py:184-274), decided on the device (K13, ``lt_vel_curriculum``).

``ModifyVelCommandsRangeBasedonReward(cfg, env)(env, env_ids, **params)``: the masked capture of episode lengths / reward sums
for the reset envs, the ``torch.all`` / ``torch.mean`` tests, the range expansion and the ``set_ranges`` bookkeeping all run in
one single-block launch that rewrites the command term's device state block -- the reference reads three reductions on the host
per branch at every reset.  Counters (``lin_forward_bins`` ...) are properties that fetch that block on demand.
``env_ids`` may be the reference's id sequence or a bool / uint8 device mask.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from .. import _C
from .._C import check, count_launches, current_stream, lib


class ModifyVelCommandsRangeBasedonReward:
    def __init__(self, cfg, env):
        lib()
        self.cfg, self._env = cfg, env
        p = cfg.params
        self.current_command = env.command_manager.get_term(p["command_name"])
        self.current_command_ranges = self.current_command.cfg.ranges
        self.command_maximum_ranges = list(p["command_maximum_ranges"])
        bins = p["curriculum_bins"]
        r = self.current_command_ranges
        self.lin_vel_x_expansion = (self.command_maximum_ranges[0] - r.lin_vel_x[1]) / bins[0]  # :191-193
        self.lin_vel_y_expansion = (self.command_maximum_ranges[1] - r.lin_vel_y[1]) / bins[1]
        self.ang_vel_z_expansion = (self.command_maximum_ranges[2] - r.ang_vel_z[1]) / bins[2]
        self.reset_envs_episode_length = p["reset_envs_episode_length"] * env.max_episode_length_s  # :194
        self.reward_name_lin, self.reward_name_ang = p["reward_name_lin"], p["reward_name_ang"]
        lin_cfg, ang_cfg = env.reward_manager.get_term_cfg(self.reward_name_lin), env.reward_manager.get_term_cfg(self.reward_name_ang)
        self.reward_threshold_lin = math.exp(-p["error_threshold_lin"] / lin_cfg.params["sigma"]) * lin_cfg.weight * env.max_episode_length_s
        self.reward_threshold_ang = math.exp(-p["error_threshold_ang"] / ang_cfg.params["sigma"]) * ang_cfg.weight * env.max_episode_length_s
        self.repeat_times_lin, self.repeat_times_ang = int(p["repeat_times_lin"]), int(p["repeat_times_ang"])
        self.max_distance_bins = int(p["max_distance_bins"])
        n, dev = env.num_envs, torch.device(env.device)
        if dev.type != "cuda":
            raise _C.LocoTouchLibraryError(f"the curriculum term needs a CUDA env (got {dev}); locotouch_b200 has no CPU path")
        self.env_num = n
        z = lambda dtype: torch.zeros(n, device=dev, dtype=dtype)  # noqa: E731
        self.env_reseted_lin, self.episode_length_buf_lin, self.episode_reward_sum_lin = z(torch.bool), z(torch.float), z(torch.float)
        self.env_reseted_ang, self.episode_length_buf_ang, self.episode_reward_sum_ang = z(torch.bool), z(torch.float), z(torch.float)

    # counters live in the command term's device block
    def _state(self):
        return self.current_command.read_state()

    lin_forward_bins = property(lambda self: int(self._state().lin_forward_bins))
    ang_forward_bins = property(lambda self: int(self._state().ang_forward_bins))
    success_repeat_times_lin = property(lambda self: int(self._state().success_repeat_times_lin))
    success_repeat_times_ang = property(lambda self: int(self._state().success_repeat_times_ang))

    def reset(self, env_ids=None):
        pass

    def __call__(self, env, env_ids, **_params):
        cmd = self.current_command
        cmd._set_mask(env_ids)
        sums = env.reward_manager._episode_sums
        a = _C.LtVelCurriculumArgs()
        a.N = self.env_num
        a.repeat_times_lin, a.repeat_times_ang, a.max_distance_bins = self.repeat_times_lin, self.repeat_times_ang, self.max_distance_bins
        a.ranges = _C.ptr(cmd._state)
        a.reset_mask = _C.ptr(cmd._mask)
        a.episode_length_buf = _C.ptr(env.episode_length_buf, torch.int64, "episode_length_buf")
        a.episode_sums_lin = _C.ptr(sums[self.reward_name_lin], torch.float32, "episode_sums[lin]")
        a.episode_sums_ang = _C.ptr(sums[self.reward_name_ang], torch.float32, "episode_sums[ang]")
        a.env_reseted_lin, a.episode_length_buf_lin, a.episode_reward_sum_lin = (
            _C.ptr(self.env_reseted_lin), _C.ptr(self.episode_length_buf_lin), _C.ptr(self.episode_reward_sum_lin))
        a.env_reseted_ang, a.episode_length_buf_ang, a.episode_reward_sum_ang = (
            _C.ptr(self.env_reseted_ang), _C.ptr(self.episode_length_buf_ang), _C.ptr(self.episode_reward_sum_ang))
        for d in range(3):
            a.command_maximum_ranges[d] = float(self.command_maximum_ranges[d])
        a.expansion[0], a.expansion[1], a.expansion[2] = self.lin_vel_x_expansion, self.lin_vel_y_expansion, self.ang_vel_z_expansion
        a.reset_envs_episode_length = float(self.reset_envs_episode_length)
        a.reward_threshold_lin, a.reward_threshold_ang = float(self.reward_threshold_lin), float(self.reward_threshold_ang)
        check(lib().lt_vel_curriculum(C.byref(a), current_stream()), "lt_vel_curriculum")
        count_launches()
