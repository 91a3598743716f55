"""Host side of K1: binds an environment's state tensors to ``lt_mdp_step`` and owns the term state.

``FusedMdp`` is what replaces, for one task, the three IsaacLab manager loops (TerminationManager.compute ->
RewardManager.compute -> [reset] -> ObservationManager.compute) and every per-term callable of
reference ``locotouch/mdp/{rewards,terminations,observations}.py`` that those loops invoke.  It reads the *same*
tensors the reference terms read (``env.scene[...].data.*``, ``env.scene.sensors[...].data.*``,
``env.command_manager.get_command``, ``env.action_manager.get_term``) by raw pointer -- no repacking.

The per-term drop-in callables in ``rewards.py`` / ``terminations.py`` / ``observations.py`` of this package are
views into the buffers filled here (SURVEY.md 8b "Fusion under a per-term API").
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from .. import _C
from .._C import check, count_launches, current_stream, lib
from . import task_spec as TS

_OBS_KIND = {
    "velocity_commands": 0,
    "base_ang_vel": 1,
    "projected_gravity": 2,
    "joint_pos": 3,
    "joint_vel": 4,
    "last_action": 5,
    "object_state": 6,
}


def _f32(x: float) -> float:
    return float(torch.tensor(x, dtype=torch.float32))


class FusedMdp:
    """One task's fused termination / reward / observation pass on ``env.device`` (must be CUDA)."""

    def __init__(self, env, spec: TS.TaskSpec, seed: int = 0):
        self.env = env
        self.spec = spec
        self.device = torch.device(env.device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise _C.LocoTouchLibraryError(f"FusedMdp needs a CUDA env (got {self.device}); locotouch_b200 has no CPU path")
        lib()  # fail loudly right away if the extension is missing
        n = self.N = env.num_envs
        dev = self.device
        nt = self.num_terms = len(spec.rewards)
        self.term_names = [t.name for t in spec.rewards]
        self.termination_names = [t.name for t in spec.terminations]
        # ---- term state (reference rewards.py:96-105) and manager buffers ([IL] RewardManager / TerminationManager)
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, device=dev, dtype=dtype)  # noqa: E731
        self.last_step_current_air_time = z(n, 4)
        self.last_step_current_contact_time = z(n, 4)
        self.swinging_in_zero_cmd = z(n, 4, dtype=torch.bool)
        self.valid_last_air_time = z(n, 4)
        self.valid_previous_contact = z(n, 4, dtype=torch.bool)
        self.last_velocity_cmd = z(n, 3)
        self.step_from_changing_cmd = z(n)
        self.reward_buf = z(n)
        self.step_reward = z(n, nt)
        self.episode_sums = z(nt, n)  # row i == RewardManager._episode_sums[name_i]
        self.term_raw = z(nt, n)
        self.term_masks = z(len(spec.terminations), n, dtype=torch.bool)
        self.terminated = z(n, dtype=torch.bool)
        self.time_outs = z(n, dtype=torch.bool)
        self.dones = z(n, dtype=torch.bool)
        self.episode_log_sums = z(nt + 1)
        d = spec.obs_dim
        self.policy_obs = z(n, d)
        self.critic_obs = z(n, d)
        self.obs_needs_fill = torch.ones(n, device=dev, dtype=torch.uint8)  # empty history: first push fills all slots [IL]
        self._any_flag = torch.full((2,), -1, device=dev, dtype=torch.int32)
        self._published_step = None  # step index for which the obs pass has published any(non_zero_cmd)
        self.seed = seed
        self.step_index = 0
        self.exact_any_nonzero_cmd = True
        self._args = _C.LtMdpArgs()
        self._bound_ptrs = None
        self._fill_static()

    # ------------------------------------------------------------------------------------------------ static part
    def _fill_static(self):
        a, spec, env = self._args, self.spec, self.env
        a.N = self.N
        a.step_dt = env.step_dt
        a.max_episode_length = int(env.max_episode_length)
        sensor = env.scene.sensors["robot_contact_senosr"]
        robot = env.scene["robot"]
        a.J = len(robot.joint_names) if hasattr(robot, "joint_names") and robot.joint_names else robot.data.joint_pos.shape[1]
        a.num_bodies = robot.data.body_pos_w.shape[1]
        feet_body = robot.find_bodies(list(spec.feet_names))[0]
        feet_sensor = sensor.find_bodies(list(spec.feet_names))[0]
        tc = sensor.find_bodies(list(spec.thigh_calf_names))[0]
        if len(feet_body) != 4 or len(feet_sensor) != 4 or len(tc) > 8:
            raise ValueError("expected 4 feet and at most 8 thigh/calf bodies")
        for k in range(4):
            a.feet_body_ids[k] = feet_body[k]
            a.feet_sensor_ids[k] = feet_sensor[k]
        for k, i in enumerate(tc):
            a.thigh_calf_sensor_ids[k] = i
        a.num_thigh_calf = len(tc)
        a.force_history = sensor.data.net_forces_w_history.shape[1]
        a.num_sensor_bodies = sensor.data.net_forces_w_history.shape[2]
        # reward table
        if len(spec.rewards) > _C.LT_MAX_REWARD_TERMS:
            raise ValueError("too many reward terms")
        a.num_reward_terms = len(spec.rewards)
        for i, t in enumerate(spec.rewards):
            a.reward_terms[i].kind = t.kind
            a.reward_terms[i].weight = t.weight
            for k, v in enumerate(t.p):
                a.reward_terms[i].p[k] = v
        a.num_termination_terms = len(spec.terminations)
        for i, t in enumerate(spec.terminations):
            tt = a.termination_terms[i]
            tt.kind, tt.time_out = t.kind, int(t.time_out)
            for k, v in enumerate(t.p):
                tt.p[k] = v
            if t.body_names is not None:
                ids = sensor.find_bodies(t.body_names)[0]
                if len(ids) > _C.LT_MAX_CONTACT_IDS:
                    raise ValueError("too many bodies in an illegal_contact term")
                tt.num_ids = len(ids)
                for k, b in enumerate(ids):
                    tt.body_ids[k] = b
        # gait parameters (reference rewards.py:61-92)
        gp, g = spec.gait, a.gait
        pair0 = sensor.find_bodies(list(gp.synced_feet_pair_names[0]))[0]
        pair1 = sensor.find_bodies(list(gp.synced_feet_pair_names[1]))[0]
        if len(gp.synced_feet_pair_names) != 2 or len(pair0) != 2 or len(pair1) != 2:
            raise ValueError("This reward only supports gaits with two pairs of synchronized feet, like trotting.")
        for k, b in enumerate([pair0[0], pair0[1], pair1[0], pair1[1]]):
            g.feet_ids[k] = b
        g.judge_time_threshold = gp.judge_time_threshold
        g.air_time_gait_bound = gp.air_time_gait_bound
        g.contact_time_gait_bound = gp.contact_time_gait_bound
        g.async_time_tolerance = gp.async_time_tolerance
        g.async_judge_time_threshold = gp.judge_time_threshold + gp.async_time_tolerance
        g.stance_rwd_scale = gp.stance_rwd_scale
        g.tolerance_proportion = gp.tolerance_proportion
        g.rwd_upper_bound = gp.rwd_upper_bound
        g.rwd_lower_bound = gp.rwd_lower_bound
        g.vel_tracking_exp_sigma = gp.vel_tracking_exp_sigma
        g.task_performance_ratio = gp.task_performance_ratio
        g.linear_scale = gp.rwd_upper_bound / (1.0 / (gp.soft_minimum_frequency * 2.0))
        g.two_step_dt = 2 * env.step_dt
        g.encourage_symmetricity = int(gp.encourage_symmetricity_and_low_frequency > 0.5)
        g.with_object = int(gp.with_object)
        g.obj_x_max, g.obj_y_max = gp.obj_x_max, gp.obj_y_max
        gs = a.gait_state
        gs.last_step_current_air_time = self.last_step_current_air_time.data_ptr()
        gs.last_step_current_contact_time = self.last_step_current_contact_time.data_ptr()
        gs.swinging_in_zero_cmd = self.swinging_in_zero_cmd.data_ptr()
        gs.valid_last_air_time = self.valid_last_air_time.data_ptr()
        gs.valid_previous_contact = self.valid_previous_contact.data_ptr()
        gs.last_velocity_cmd = self.last_velocity_cmd.data_ptr()
        gs.step_from_changing_cmd = self.step_from_changing_cmd.data_ptr()
        # outputs
        a.reward = self.reward_buf.data_ptr()
        a.step_reward = self.step_reward.data_ptr()
        a.episode_sums = self.episode_sums.data_ptr()
        a.term_raw = self.term_raw.data_ptr()
        a.term_masks = self.term_masks.data_ptr()
        a.terminated = self.terminated.data_ptr()
        a.time_outs = self.time_outs.data_ptr()
        a.dones = self.dones.data_ptr()
        a.episode_log_sums = self.episode_log_sums.data_ptr()
        a.any_flag_ws = self._any_flag.data_ptr()
        # observations
        a.num_obs_terms = len(spec.obs_terms)
        a.history_length = spec.history_length
        for i, t in enumerate(spec.obs_terms):
            ot = a.obs_terms[i]
            ot.kind, ot.dim, ot.scale = _OBS_KIND[t.name], t.dim, t.scale
            ot.noisy = int(t.noise is not None)
            ot.n_min, ot.n_max = t.noise if t.noise is not None else (0.0, 0.0)
        if spec.object_state is not None:
            os_ = spec.object_state
            # the cfg lists 12 noise bounds: pos(3) vel(3) EULER(3) ang-vel(3); the quaternion slots get no additive noise
            n_min = tuple(os_.n_min[0:6]) + (0.0,) * 4 + tuple(os_.n_min[9:])
            n_max = tuple(os_.n_max[0:6]) + (0.0,) * 4 + tuple(os_.n_max[9:])
            for k in range(13):
                a.os_n_min[k], a.os_n_max[k] = n_min[k], n_max[k]
                a.os_scale[k], a.os_non_contact[k] = os_.scale[k], os_.non_contact_obs[k]
            for k in range(3):
                a.os_euler_min[k], a.os_euler_max[k] = os_.n_min[6 + k], os_.n_max[6 + k]
            a.os_last_contact_thr = os_.last_contact_time_threshold
            a.os_current_contact_thr = os_.current_contact_time_threshold
        self._build_tables()

    def _build_tables(self):
        """Launch-constant lookup tables (lt_mdp_build_tables), computed once on the host and kept on the device; must be
        rebuilt whenever the term tables change (kinds, dims, zero weights)."""
        a = self._args
        n = lib().lt_mdp_tables_len(C.byref(a))
        if n <= 0:
            raise _C.LocoTouchLibraryError("lt_mdp_tables_len: invalid term tables")
        host = (C.c_int32 * n)()
        check(lib().lt_mdp_build_tables(C.byref(a), host, n), "lt_mdp_build_tables")
        self._tables = torch.tensor(list(host), dtype=torch.int32).to(self.device)
        a.tables = self._tables.data_ptr()

    # --------------------------------------------------------------------------------------- per-step pointer binding
    def _tensor_fields(self):
        env = self.env
        r = env.scene["robot"].data
        s = env.scene.sensors["robot_contact_senosr"].data
        t = env.action_manager.get_term("joint_pos")
        f = {
            "command": env.command_manager.get_command("base_velocity"),
            "root_pos_w": r.root_pos_w, "root_lin_vel_b": r.root_lin_vel_b, "root_ang_vel_b": r.root_ang_vel_b,
            "projected_gravity_b": r.projected_gravity_b, "joint_pos": r.joint_pos, "joint_vel": r.joint_vel,
            "joint_acc": r.joint_acc, "applied_torque": r.applied_torque, "default_joint_pos": r.default_joint_pos,
            "default_joint_vel": r.default_joint_vel, "soft_joint_pos_limits": r.soft_joint_pos_limits,
            "raw_actions": t.raw_actions, "prev_raw_actions": t.prev_raw_actions,
            "body_pos_w": r.body_pos_w, "body_lin_vel_w": r.body_lin_vel_w,
            "net_forces_w_history": s.net_forces_w_history, "current_air_time": s.current_air_time,
            "current_contact_time": s.current_contact_time, "last_air_time": s.last_air_time,
            "episode_length_buf": env.episode_length_buf,
        }
        if self.spec.with_object:
            o = env.scene["object"].data
            oc = env.scene.sensors["object_contact_sensor"].data
            f.update({
                "root_quat_w": r.root_quat_w, "root_lin_vel_w": r.root_lin_vel_w, "root_ang_vel_w": r.root_ang_vel_w,
                "obj_root_pos_w": o.root_pos_w, "obj_root_quat_w": o.root_quat_w, "obj_root_lin_vel_w": o.root_lin_vel_w,
                "obj_root_ang_vel_w": o.root_ang_vel_w, "obj_projected_gravity_b": o.projected_gravity_b,
                "obj_last_contact_time": oc.last_contact_time, "obj_current_contact_time": oc.current_contact_time,
                "obj_current_air_time": oc.current_air_time,
            })
        return f

    def bind(self):
        """(Re)reads the data pointers of the env tensors.  Cheap; called before every launch."""
        fields = self._tensor_fields()
        ptrs = tuple(t.data_ptr() for t in fields.values())
        if ptrs == self._bound_ptrs:
            return
        a = self._args
        for name, t in fields.items():
            want = torch.int64 if name == "episode_length_buf" else torch.float32
            if t.dtype != want or not t.is_contiguous() or t.device != self.device:
                raise _C.LocoTouchLibraryError(
                    f"env tensor {name} must be a contiguous {want} tensor on {self.device} "
                    f"(got {t.dtype}, {t.device}, contiguous={t.is_contiguous()})")
            setattr(a, name, t.data_ptr())
        self._keepalive = fields
        self._bound_ptrs = ptrs

    # --------------------------------------------------------------------------------------------------------- launch
    def step(self, rewards: bool = True, observations: bool = True, *, auto_reset: bool = True, u_obs=None, u_obj_euler=None,
             policy_out=None, critic_out=None, policy_in=None, critic_in=None, any_nonzero_cmd: bool | None = None,
             step_offset: int | None = None, offset_base: torch.Tensor | None = None, actions=None, action_term=None,
             reset_action_term: bool = False, store=None):
        """One fused pass.  ``rewards``: terminations + rewards (+ reset of done envs when ``auto_reset``);
        ``observations``: policy / critic observation rows (history source ``*_in`` defaults to this object's buffers,
        destination ``*_out`` likewise; they may alias, or point into RolloutStorage slots for a zero-copy rollout).
        ``offset_base`` (device int64 scalar) + ``step_offset``: the env-step index lives on the device, so that the call can
        be captured in a CUDA graph and replayed (fresh noise, correct any(non_zero_cmd) hand-over) -- the caller then
        guarantees that the previous step's observation pass ran with the same counter."""
        self.bind()
        a = self._args
        # ``actions`` + ``action_term`` (dict: prev_prev_raw, processed, offset tensors or None; clip, raw_scale, scale): the action term's
        # process_actions (reference mdp/actions.py:30-44 = K0) runs INSIDE this launch on the bound raw / prev_raw tensors, before any
        # term reads them -- one launch less per env step for a caller that owns the step loop
        if actions is not None:
            if not rewards:
                raise _C.LocoTouchLibraryError("FusedMdp.step(actions=...): the fused action term needs the reward pass")
            t = action_term or {}
            a.act_new = _C.ptr(actions, torch.float32, "actions")
            a.act_prev_prev_raw = _C.ptr(t.get("prev_prev_raw"), torch.float32)
            a.act_processed = _C.ptr(t.get("processed"), torch.float32)
            a.act_offset = _C.ptr(t.get("offset"), torch.float32)
            a.act_clip, a.act_raw_scale, a.act_scale = float(t.get("clip", 0.0)), float(t.get("raw_scale", 1.0)), float(t.get("scale", 1.0))
        else:
            a.act_new = None
            a.act_prev_prev_raw = _C.ptr((action_term or {}).get("prev_prev_raw"), torch.float32)
        # ``reset_action_term``: IsaacLab's ActionManager.reset(env_ids) (reference mdp/actions.py:46-52) for the envs this launch resets,
        # between the reward and the observation pass as ManagerBasedRLEnv.step orders them: raw / prev_raw (/ prev_prev_raw when given
        # in ``action_term``) zeroed in place, the post-reset last_action observation built from the zeroed row
        a.act_reset_on_done = int(bool(reset_action_term and auto_reset and rewards))
        # ``store`` (dict: rewards [N] float slot, dones [N] uint8 slot or None, values [N] or None, gamma): K3 inside this launch -- the
        # time-out bootstrap of PPO.process_env_step and the scalar rollout store (reference ppo.py:162-165, rollout_storage.py:86-88)
        # written straight into the RolloutStorage row of this env step
        if store is not None:
            if not rewards:
                raise _C.LocoTouchLibraryError("FusedMdp.step(store=...): the fused rollout store needs the reward pass")
            a.store_rewards = _C.ptr(store["rewards"], torch.float32, "store rewards")
            a.store_dones = _C.ptr(store.get("dones"), torch.uint8)
            a.store_values = _C.ptr(store.get("values"), torch.float32)
            a.store_gamma = float(store.get("gamma", 0.0))
            for k in ("rewards", "dones", "values"):
                if store.get(k) is not None and store[k].numel() != self.N:
                    raise _C.LocoTouchLibraryError(f"FusedMdp.step(store=...): {k} must hold one element per env")
        else:
            a.store_rewards = None
        a.phases = (_C.LT_PHASE_REWARDS if rewards else 0) | (_C.LT_PHASE_OBS if observations else 0)
        a.auto_reset = int(auto_reset)
        # an observation-only pass belongs to the step whose reward pass already ran (IsaacLab order: rewards -> reset -> obs)
        step = self.step_index if rewards else self.step_index - 1
        if offset_base is not None:
            step = int(step_offset or 0)
            a.offset_base = _C.ptr(offset_base, torch.int64, "offset_base")
        else:
            a.offset_base = None
        a.seed, a.offset = self.seed, step & 0xFFFFFFFFFFFFFFFF
        if offset_base is not None and any_nonzero_cmd is None:
            a.any_nonzero_cmd_override = -1
        elif any_nonzero_cmd is not None:
            a.any_nonzero_cmd_override = int(bool(any_nonzero_cmd))
        elif not self.exact_any_nonzero_cmd:
            a.any_nonzero_cmd_override = 1
        elif self._published_step == self.step_index:
            a.any_nonzero_cmd_override = -1  # the previous observation pass already reduced it on the device
        else:
            a.any_nonzero_cmd_override = -2  # run the one-block reduction first
        n_launch = 1 + (1 if (rewards and a.any_nonzero_cmd_override == -2) else 0)
        a.obs_fill = self.obs_needs_fill.data_ptr()
        if observations:
            pin = self.policy_obs if policy_in is None else policy_in
            cin = self.critic_obs if critic_in is None else critic_in
            pout = self.policy_obs if policy_out is None else policy_out
            cout = self.critic_obs if critic_out is None else critic_out
            a.policy_obs_in, a.critic_obs_in = _C.ptr(pin, torch.float32), _C.ptr(cin, torch.float32)
            a.policy_obs_out, a.critic_obs_out = _C.ptr(pout, torch.float32), _C.ptr(cout, torch.float32)
            a.u_obs = _C.ptr(u_obs, torch.float32, "u_obs")
            a.u_obj_euler = _C.ptr(u_obj_euler, torch.float32, "u_obj_euler")
        check(lib().lt_mdp_step(C.byref(a), current_stream()), "lt_mdp_step")
        count_launches(n_launch)
        if offset_base is None:
            if observations:
                self._published_step = step + 1
            if rewards:
                self.step_index += 1
        return self

    def compute_rewards(self, **kw):
        """terminations + rewards only (what IsaacLab runs before it resets the done envs)."""
        return self.step(True, False, **kw)

    def compute_observations(self, **kw):
        """observations only (what IsaacLab runs after reset + command resampling)."""
        return self.step(False, True, **kw)

    def reset(self, env_ids=None):
        """``ManagerTermBase.reset(env_ids)`` of the gait term + RewardManager.reset: zero the state of those envs."""
        mask = torch.zeros(self.N, device=self.device, dtype=torch.bool)
        if env_ids is None:
            mask[:] = True
        else:
            mask[env_ids] = True
        gs = self._args.gait_state
        check(lib().lt_mdp_reset(C.byref(gs), self.episode_sums.data_ptr(), self.num_terms, mask.view(torch.uint8).data_ptr(), self.N,
                                 current_stream()), "lt_mdp_reset")
        count_launches(1)
        self.obs_needs_fill |= mask.to(torch.uint8)

    # ------------------------------------------------------------------------------------------------------- accessors
    def term(self, name: str) -> torch.Tensor:
        return self.term_raw[self.term_names.index(name)]

    def termination(self, name: str) -> torch.Tensor:
        return self.term_masks[self.termination_names.index(name)]

    def gait_state_dict(self):
        return dict(
            last_step_current_air_time=self.last_step_current_air_time,
            last_step_current_contact_time=self.last_step_current_contact_time,
            swinging_in_zero_cmd=self.swinging_in_zero_cmd,
            valid_last_air_time=self.valid_last_air_time,
            valid_previous_contact=self.valid_previous_contact,
            last_velocity_cmd=self.last_velocity_cmd,
            step_from_changing_cmd=self.step_from_changing_cmd,
        )
