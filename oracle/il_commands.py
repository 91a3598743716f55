"""ORACLE (test infrastructure, never imported by the product path).

[IL] Restatement of the two IsaacLab 2.2.1 classes the LocoTouch command terms derive from --
``isaaclab.managers.CommandTerm`` and ``isaaclab.envs.mdp.commands.UniformVelocityCommand`` (+ their cfg classes).
IsaacLab is not under /root/reference and not installable here, so these are written from its published behaviour
(SURVEY.md App. B): parity for THIS file is UNPINNED.  Its only purpose is to let the UNMODIFIED reference subclasses
(``locotouch/mdp/commands.py:379-595``) run in the build container (``oracle/ref_loader.load_reference_commands``) so that
the restatement in ``oracle/commands.py`` can be pinned against them.  Debug visualisation is left out.
"""
from __future__ import annotations

import torch


class CommandTermCfg:
    class_type = None
    resampling_time_range = (0.0, 0.0)
    debug_vis = False

    def __init__(self, **kw):
        for klass in reversed(type(self).__mro__):
            for key, val in vars(klass).items():
                if not key.startswith("__") and not callable(val):
                    setattr(self, key, val)
        self.__dict__.update(kw)


class UniformVelocityCommandCfg(CommandTermCfg):
    asset_name = "robot"
    heading_command = False
    heading_control_stiffness = 1.0
    rel_standing_envs = 0.0
    rel_heading_envs = 1.0

    class Ranges:
        def __init__(self, lin_vel_x=(0.0, 0.0), lin_vel_y=(0.0, 0.0), ang_vel_z=(0.0, 0.0), heading=None):
            self.lin_vel_x, self.lin_vel_y, self.ang_vel_z, self.heading = lin_vel_x, lin_vel_y, ang_vel_z, heading

    ranges = None


class CommandTerm:
    """[IL] isaaclab.managers.CommandTerm: resampling timer, per-env metrics, reset / compute drivers."""

    def __init__(self, cfg, env):
        self.cfg, self._env = cfg, env
        self.metrics = dict()
        self.time_left = torch.zeros(self.num_envs, device=self.device)
        self.command_counter = torch.zeros(self.num_envs, device=self.device, dtype=torch.long)

    @property
    def num_envs(self):
        return self._env.num_envs

    @property
    def device(self):
        return self._env.device

    def reset(self, env_ids=None):
        if env_ids is None:
            env_ids = slice(None)
        extras = {}
        for name, value in self.metrics.items():
            extras[name] = torch.mean(value[env_ids]).item()  # mean over the reset envs, then cleared for them
            value[env_ids] = 0.0
        self.command_counter[env_ids] = 0
        self._resample(env_ids)
        return extras

    def compute(self, dt: float):
        self._update_metrics()
        self.time_left -= dt
        resample_env_ids = (self.time_left <= 0.0).nonzero().flatten()
        if len(resample_env_ids) > 0:
            self._resample(resample_env_ids)
        self._update_command()

    def _resample(self, env_ids):
        if isinstance(env_ids, slice):
            env_ids = torch.arange(self.num_envs, device=self.device)[env_ids]
        if len(env_ids) != 0:
            self.time_left[env_ids] = self.time_left[env_ids].uniform_(*self.cfg.resampling_time_range)
            self._resample_command(env_ids)
            self.command_counter[env_ids] += 1


class UniformVelocityCommand(CommandTerm):
    """[IL] isaaclab.envs.mdp.commands.UniformVelocityCommand (heading control included for completeness)."""

    def __init__(self, cfg, env):
        super().__init__(cfg, env)
        if self.cfg.heading_command and self.cfg.ranges.heading is None:
            raise ValueError("The velocity command has heading commands active (heading_command=True) but the `ranges.heading` parameter is set to None.")
        self.robot = env.scene[cfg.asset_name]
        self.vel_command_b = torch.zeros(self.num_envs, 3, device=self.device)
        self.heading_target = torch.zeros(self.num_envs, device=self.device)
        self.is_heading_env = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self.is_standing_env = torch.zeros_like(self.is_heading_env)
        self.metrics["error_vel_xy"] = torch.zeros(self.num_envs, device=self.device)
        self.metrics["error_vel_yaw"] = torch.zeros(self.num_envs, device=self.device)

    @property
    def command(self):
        return self.vel_command_b

    def _update_metrics(self):
        max_command_step = self.cfg.resampling_time_range[1] / self._env.step_dt
        self.metrics["error_vel_xy"] += torch.norm(self.vel_command_b[:, :2] - self.robot.data.root_lin_vel_b[:, :2], dim=-1) / max_command_step
        self.metrics["error_vel_yaw"] += torch.abs(self.vel_command_b[:, 2] - self.robot.data.root_ang_vel_b[:, 2]) / max_command_step

    def _resample_command(self, env_ids):
        r = torch.empty(len(env_ids), device=self.device)
        self.vel_command_b[env_ids, 0] = r.uniform_(*self.cfg.ranges.lin_vel_x)
        self.vel_command_b[env_ids, 1] = r.uniform_(*self.cfg.ranges.lin_vel_y)
        self.vel_command_b[env_ids, 2] = r.uniform_(*self.cfg.ranges.ang_vel_z)
        if self.cfg.heading_command:
            self.heading_target[env_ids] = r.uniform_(*self.cfg.ranges.heading)
            self.is_heading_env[env_ids] = r.uniform_(0.0, 1.0) <= self.cfg.rel_heading_envs
        self.is_standing_env[env_ids] = r.uniform_(0.0, 1.0) <= self.cfg.rel_standing_envs

    def _update_command(self):
        if self.cfg.heading_command:
            env_ids = self.is_heading_env.nonzero(as_tuple=False).flatten()
            err = self.heading_target[env_ids] - self.robot.data.heading_w[env_ids]
            err = (err + torch.pi) % (2 * torch.pi) - torch.pi
            self.vel_command_b[env_ids, 2] = torch.clip(self.cfg.heading_control_stiffness * err, min=self.cfg.ranges.ang_vel_z[0],
                                                        max=self.cfg.ranges.ang_vel_z[1])
        standing_env_ids = self.is_standing_env.nonzero(as_tuple=False).flatten()
        self.vel_command_b[standing_env_ids, :] = 0.0
