"""Observation terms with the reference's names / signatures (reference locotouch/mdp/observations.py).

* ``object_state_in_robot_frame``  observations.py:38-91  -> 13-D rows of a history-1 fused observation pass (K1)
* ``BinaryTactileSignals``         observations.py:95-159,166-199,281-308 -> K2
The other ``TactileSignals`` encodings are "next" scope.
"""
from __future__ import annotations

import copy

import torch

from .. import ops
from . import task_spec as TS
from ._fusion import cache_for
from .fused import FusedMdp

__all__ = ["object_state_in_robot_frame", "BinaryTactileSignals"]


def object_state_in_robot_frame(env, robot_cfg=None, object_cfg=None, sensor_cfg=None, last_contact_time_threshold: float = 0.00001,
                                current_contact_time_threshold: float = 0.00001, non_contact_obs: list = [0.0] * 13,
                                add_uniform_noise: bool = False, n_min=-0.03, n_max=0.03, scale=1.0) -> torch.Tensor:
    c = cache_for(env)
    step = getattr(env, "common_step_counter", None)
    if c.obj_obs is None:
        base = c.fused.spec
        os_ = TS.ObjectStateObs(
            n_min=tuple(n_min) if not isinstance(n_min, float) else (n_min,) * 12, n_max=tuple(n_max) if not isinstance(n_max, float) else (n_max,) * 12,
            scale=tuple(scale) if not isinstance(scale, float) else (scale,) * 13, non_contact_obs=tuple(non_contact_obs),
            last_contact_time_threshold=last_contact_time_threshold, current_contact_time_threshold=current_contact_time_threshold)
        spec = copy.deepcopy(base)
        spec.obs_terms = [TS.ObsTerm("object_state", 13, 1.0, None)]
        spec.history_length = 1
        spec.object_state = os_
        c.obj_obs = FusedMdp(env, spec, seed=base.max_episode_length)
        c.obj_step = None
    f = c.obj_obs
    if c.obj_step is None or step is None or step != c.obj_step:
        f.env = env
        f.compute_observations()
        c.obj_step = step
    return (f.policy_obs if add_uniform_noise else f.critic_obs).clone()


class BinaryTactileSignals:
    """Class term: per-(env, taxel) thresholds sampled once at construction (observations.py:121-126), every call is one
    ``lt_taxel_synth`` launch (in-kernel Philox dropout / addition).  Side buffers keep the reference's attribute names."""

    def __init__(self, cfg, env):
        self.cfg = cfg
        self._env = env
        p = cfg.params
        self.asset_cfg, self.sensor_cfg = p.get("asset_cfg"), p.get("sensor_cfg")
        self.asset = env.scene[self.asset_cfg.name]
        self.contact_sensor = env.scene.sensors[self.sensor_cfg.name]
        rows, cols = p.get("tactile_signal_shape")
        self.tactile_signals_shape = (env.num_envs, rows, cols)
        dev = env.device
        n, t = env.num_envs, rows * cols
        self.contact_threshold = p.get("contact_threshold")
        thr = torch.ones(self.tactile_signals_shape, device=dev) * self.contact_threshold
        if p.get("add_threshold_noise"):
            lo, hi = p.get("threshold_n_min"), p.get("threshold_n_max")
            thr = self.contact_threshold + torch.rand_like(thr) * (hi - lo) + lo
        self.contact_threshold_envs_sensors = thr.contiguous()
        self.contact_dropout_prob = float(p.get("contact_dropout_prob"))
        self.contact_addition_prob = float(p.get("contact_addition_prob"))
        if p.get("add_continuous_artifact", 0.0) > 0.5:
            raise NotImplementedError("add_continuous_artifact is unreachable in the reference (undefined N at observations.py:136)")
        self.original_contact_taxels = torch.zeros(n, rows, cols, device=dev, dtype=torch.bool)
        self.original_normal_forces = torch.zeros(n, rows, cols, device=dev)
        self.processed_contact_taxels = torch.zeros(n, rows, cols, device=dev, dtype=torch.bool)
        self._signal = torch.zeros(n, 2 * t, device=dev)
        self._packed = torch.zeros(n, (t + 31) // 32, device=dev, dtype=torch.int32)
        ids = self.asset_cfg.body_ids
        self._offset = 0 if isinstance(ids, slice) else int(ids[0])
        if not isinstance(ids, slice) and list(ids) != list(range(ids[0], ids[0] + t)):
            raise ValueError("taxel links must be a contiguous block of the articulation's bodies")
        self.seed, self._calls = 0, 0

    num_envs = property(lambda self: self._env.num_envs)
    device = property(lambda self: self._env.device)

    def reset(self, env_ids=None):
        pass

    def __call__(self, env, u_drop=None, u_add=None, **params) -> torch.Tensor:
        n, rows, cols = self.tactile_signals_shape
        t = rows * cols
        forces = self.contact_sensor.data.net_forces_w
        ids = self.sensor_cfg.body_ids
        if not isinstance(ids, slice):
            forces = forces[:, ids].contiguous()
        ops.taxel_synth(self.asset.data.body_quat_w, forces, self.contact_threshold_envs_sensors.view(n, t), quat_body_offset=self._offset,
                        u_drop=u_drop, u_add=u_add, p_drop=self.contact_dropout_prob, p_add=self.contact_addition_prob, seed=self.seed,
                        offset=self._calls, signal=self._signal, packed=self._packed, normal_forces=self.original_normal_forces.view(n, t),
                        original_contact=self.original_contact_taxels.view(n, t))
        self._calls += 1
        self.processed_contact_taxels = self._signal[:, :t].view(n, rows, cols) > 0.5
        return self._signal.clone()


class _ForceTactileSignals(BinaryTactileSignals):
    """Shared part of the force-valued encodings (reference observations.py:166-237): one ``lt_taxel_forces`` launch per call writes
    the channels the subclass lists straight into the [N, C, T] observation; parameters are read from ``cfg.params`` at
    construction like the reference's ``TactileSignals.__init__`` (observations.py:96-152)."""

    channels: tuple = ()

    def __init__(self, cfg, env):
        super().__init__(cfg, env)
        p = cfg.params
        self.add_force_noise = bool(p.get("add_force_noise"))
        self.force_n_prop_min = float(p.get("force_n_prop_min") or 0.0)
        self.force_n_prop_max = float(p.get("force_n_prop_max") or 0.0)
        self.maximal_force = float(p.get("maximal_force"))
        self.total_levels = int(p.get("total_levels"))
        self.add_level_noise = bool(p.get("add_level_noise"))
        self.level_n_min = float(p.get("level_n_min") or 0.0)
        self.level_n_max = float(p.get("level_n_max") or 0.0)
        n, rows, cols = self.tactile_signals_shape
        self._out = torch.zeros(n, len(self.channels), rows * cols, device=env.device)

    def __call__(self, env, u=None, **params) -> torch.Tensor:
        n, rows, cols = self.tactile_signals_shape
        t = rows * cols
        forces = self.contact_sensor.data.net_forces_w
        ids = self.sensor_cfg.body_ids
        if not isinstance(ids, slice):
            forces = forces[:, ids].contiguous()
        ops.taxel_forces(self.asset.data.body_quat_w, forces, self.contact_threshold_envs_sensors.view(n, t), self._out, self.channels,
                         quat_body_offset=self._offset, u=u, p_drop=self.contact_dropout_prob, p_add=self.contact_addition_prob,
                         add_force_noise=self.add_force_noise, force_n_prop_min=self.force_n_prop_min, force_n_prop_max=self.force_n_prop_max,
                         maximal_force=self.maximal_force, total_levels=self.total_levels, add_level_noise=self.add_level_noise,
                         level_n_min=self.level_n_min, level_n_max=self.level_n_max, seed=self.seed, offset=self._calls)
        self._calls += 1
        self.processed_contact_taxels = self._out[:, 0].view(n, rows, cols) > 0.5
        return self._out.view(n, -1).clone()


class NormalizedTactileSignals(_ForceTactileSignals):
    """[contact, per-env min-max normalised force] (reference observations.py:311-337)."""
    channels = ("contact", "minmax")


class DiscreteTactileSignals(_ForceTactileSignals):
    """[contact, discretised signal] (reference observations.py:340-366)."""
    channels = ("contact", "discretized")


class CotinuousTactileSignals(_ForceTactileSignals):
    """[contact, force / maximal_force clamped to [0, 1]] (reference observations.py:369-396; the reference's spelling)."""
    channels = ("contact", "normalized")


class ProcessedTactileSignals(_ForceTactileSignals):
    """[contact, normalised, min-max normalised, discretised] (reference observations.py:399-429)."""
    channels = ("contact", "normalized", "minmax", "discretized")
