"""Velocity command term with the reference's class name and CommandTerm interface (reference locotouch/mdp/commands.py:379-576),
running on the device (K13, ``lt_command_step``).

``UniformVelocityCommandGaitLoggingMultiSampling(cfg, env)`` keeps the attributes the reference (and IsaacLab's CommandManager)
touch: ``command`` / ``vel_command_b``, ``vel_command_b_buffer``, ``time_left``, ``command_counter``, ``is_standing_env``,
``metrics[name]`` ([N] rows), ``cfg.ranges`` / ``cfg.previous_ranges`` / ``cfg.rel_standing_envs``, ``*_equal_ranges``,
``initial_zero_command_steps``, ``set_ranges(...)``, ``reset(env_ids) -> extras``, ``compute(dt)``.

What differs, by design: the ranges and the curriculum counters live in ONE device block (``LtCommandRanges``) that the curriculum
kernel rewrites and the command kernel reads, so a reset -> curriculum -> resample -> compute sequence needs no host read;
attribute reads such as ``cfg.ranges.lin_vel_x`` fetch that block on demand.  ``reset`` accepts the reference's env-id
sequence or a bool / uint8 device mask (``FusedMdp.dones``) -- the mask form has no ``nonzero``.  Random draws are Philox streams
keyed by (seed, call index, env) instead of torch's global generator; explicit uniforms can be passed for parity tests.
``heading_command`` is not supported (every LocoTouch cfg sets it False, reference locomotion_base_env_cfg.py:58).
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _C
from .._C import check, count_launches, current_stream, lib

_KEYS = ("lin_vel_x", "lin_vel_y", "ang_vel_z")


class _RangesView:
    """``cfg.ranges`` / ``cfg.previous_ranges``: attribute reads fetch the device block, writes go through ``set_ranges``."""

    def __init__(self, term, field):
        object.__setattr__(self, "_term", term)
        object.__setattr__(self, "_field", field)

    def __getattr__(self, key):
        if key not in _KEYS:
            raise AttributeError(key)
        arr = getattr(self._term.read_state(), self._field)[_KEYS.index(key)]
        return (arr[0], arr[1])

    def __setattr__(self, key, value):
        if key not in _KEYS:
            raise AttributeError(key)
        st = self._term.read_state()
        arr = getattr(st, self._field)[_KEYS.index(key)]
        arr[0], arr[1] = float(value[0]), float(value[1])
        self._term.write_state(st)


class _Metrics:
    """dict-like view of the [14, N] metrics block: ``metrics[name]`` is the [N] row (writable, like the reference's tensors)."""

    def __init__(self, block):
        self._block = block

    def __getitem__(self, name):
        return self._block[_C.LT_CMD_METRICS.index(name)]

    def __setitem__(self, name, value):
        self._block[_C.LT_CMD_METRICS.index(name)].copy_(torch.as_tensor(value, device=self._block.device))

    def keys(self):
        return list(_C.LT_CMD_METRICS)

    def items(self):
        return [(k, self[k]) for k in _C.LT_CMD_METRICS]

    def __iter__(self):
        return iter(_C.LT_CMD_METRICS)

    def __len__(self):
        return len(_C.LT_CMD_METRICS)


class UniformVelocityCommandGaitLoggingMultiSampling:
    def __init__(self, cfg, env, seed: int = 0):
        if getattr(cfg, "heading_command", False):
            raise NotImplementedError("heading_command=True is not part of the LocoTouch tasks (locomotion_base_env_cfg.py:58)")
        lib()
        self.cfg_static, self._env = cfg, env
        dev = torch.device(env.device)
        if dev.type != "cuda":
            raise _C.LocoTouchLibraryError(f"the command term needs a CUDA env (got {dev}); locotouch_b200 has no CPU path")
        self.device, n = dev, env.num_envs
        self.num_envs = n
        p = float(cfg.new_command_probs)
        if 1.0 - 2.0 * p < 0:
            raise ValueError(f"new_command_probs too large: {p}, must satisfy 1-2p >= 0.")
        # state block on the device (commands.py:431-436, 451 and the curriculum counters)
        st = _C.LtCommandRanges()
        for d, key in enumerate(_KEYS):
            lo, hi = getattr(cfg.ranges, key)
            st.ranges[d][0] = st.previous[d][0] = float(lo)
            st.ranges[d][1] = st.previous[d][1] = float(hi)
            st.equal[d] = 1
        st.initial_zero_command_steps = int(cfg.initial_zero_command_steps)
        st.final_initial_zero_command_steps = int(cfg.final_initial_zero_command_steps)
        st.rel_standing_envs = float(cfg.rel_standing_envs)
        st.final_rel_standing_envs = float(cfg.final_rel_standing_envs)
        self._state = torch.zeros(C.sizeof(st), dtype=torch.uint8, device=dev)
        self._state_host = torch.zeros(C.sizeof(st), dtype=torch.uint8).pin_memory()
        self.write_state(st)
        # what the reference keeps in cfg and mutates: views onto the device block
        self.cfg = _CfgView(self, cfg)
        probs = torch.tensor([p, 1.0 - 2 * p, p])  # :448, fp32 like the reference tensor
        cum = torch.cumsum(probs, 0) / probs.sum()
        self._bin_c = (float(cum[0]), float(cum[1]))
        self.binary_maximal_command = bool(cfg.binary_maximal_command)
        # [IL] CommandTerm / UniformVelocityCommand state
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, device=dev, dtype=dtype)  # noqa: E731
        self.vel_command_b = z(n, 3)
        self.vel_command_b_buffer = z(n, 3)
        self.time_left = z(n)
        self.command_counter = z(n, dtype=torch.long)
        self.is_standing_env = z(n, dtype=torch.bool)
        self._metrics = z(len(_C.LT_CMD_METRICS), n)
        self.metrics = _Metrics(self._metrics)
        self._metrics[_C.LT_CMD_METRICS.index("initial_zero_command_steps")] = float(cfg.initial_zero_command_steps)  # :468-469
        self._metrics[_C.LT_CMD_METRICS.index("rel_standing_envs")] = float(cfg.rel_standing_envs)
        self._scalars = z(len(_C.LT_CMD_METRICS))
        self._extras = z(len(_C.LT_CMD_METRICS) + 1, dtype=torch.float64)
        self._ws = torch.zeros(int(lib().lt_command_workspace_bytes(n)), dtype=torch.uint8, device=dev)
        self._mask = z(n, dtype=torch.uint8)
        sensor_cfg = cfg.sensor_cfg
        sensor_cfg.resolve(env.scene)  # :382-384
        self.sensor_cfg = sensor_cfg
        self.contact_sensor = env.scene.sensors[sensor_cfg.name]
        self.robot = env.scene[cfg.asset_name]
        self.seed, self.calls = seed, 0

    # ------------------------------------------------------------------------------------------ device state block
    def read_state(self) -> _C.LtCommandRanges:
        """Device -> host copy of the scalar state (a synchronising read: logging / inspection, not the step path)."""
        self._state_host.copy_(self._state)
        return _C.LtCommandRanges.from_buffer_copy(bytes(self._state_host.numpy().tobytes()))

    def write_state(self, st: _C.LtCommandRanges):
        self._state.copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))

    @property
    def command(self) -> torch.Tensor:
        return self.vel_command_b

    @property
    def lin_vel_x_equal_ranges(self) -> bool:
        return bool(self.read_state().equal[0])

    @property
    def lin_vel_y_equal_ranges(self) -> bool:
        return bool(self.read_state().equal[1])

    @property
    def ang_vel_z_equal_ranges(self) -> bool:
        return bool(self.read_state().equal[2])

    @property
    def initial_zero_command_steps(self) -> int:
        return int(self.read_state().initial_zero_command_steps)

    def set_ranges(self, lin_vel_x=None, lin_vel_y=None, ang_vel_z=None):
        """commands.py:471-505 (host-side entry; the curriculum kernel does the same on the device)."""
        st = self.read_state()
        for d, new in enumerate((lin_vel_x, lin_vel_y, ang_vel_z)):
            if new is not None:
                st.previous[d][0], st.previous[d][1] = st.ranges[d][0], st.ranges[d][1]
                st.ranges[d][0], st.ranges[d][1] = float(new[0]), float(new[1])
                st.equal[d] = int(st.previous[d][0] == st.ranges[d][0] and st.previous[d][1] == st.ranges[d][1])
        if st.equal[0] and st.equal[1] and st.equal[2]:
            st.initial_zero_command_steps = st.final_initial_zero_command_steps
            st.rel_standing_envs = st.final_rel_standing_envs
        self.write_state(st)

    # ------------------------------------------------------------------------------------------------- launches
    def _args(self, phases: int, dt: float, u) -> _C.LtCommandArgs:
        env, a = self._env, _C.LtCommandArgs()
        a.N, a.phases, a.dt = self.num_envs, phases, float(dt)
        a.resampling_time_lo, a.resampling_time_hi = (float(x) for x in self.cfg_static.resampling_time_range)
        a.bin_c0, a.bin_c1 = self._bin_c
        a.binary_maximal_command = int(self.binary_maximal_command)
        a.ranges = _C.ptr(self._state)
        a.vel_command_b = _C.ptr(self.vel_command_b, torch.float32, "vel_command_b")
        a.vel_command_b_buffer = _C.ptr(self.vel_command_b_buffer, torch.float32)
        a.time_left = _C.ptr(self.time_left, torch.float32)
        a.command_counter = _C.ptr(self.command_counter, torch.int64)
        a.is_standing_env = _C.ptr(self.is_standing_env)
        a.metrics, a.metric_scalars = _C.ptr(self._metrics), _C.ptr(self._scalars)
        a.reset_mask, a.reset_extras = _C.ptr(self._mask), _C.ptr(self._extras)
        a.episode_length_buf = _C.ptr(env.episode_length_buf, torch.int64, "episode_length_buf")
        a.root_lin_vel_b = _C.ptr(self.robot.data.root_lin_vel_b, torch.float32, "root_lin_vel_b")
        a.root_ang_vel_b = _C.ptr(self.robot.data.root_ang_vel_b, torch.float32, "root_ang_vel_b")
        last_air = self.contact_sensor.data.last_air_time
        a.last_air_time = _C.ptr(last_air, torch.float32, "last_air_time")
        a.num_sensor_bodies = last_air.shape[1]
        ids = self.sensor_cfg.body_ids
        ids = list(range(last_air.shape[1]))[ids] if isinstance(ids, slice) else list(ids)
        if len(ids) != 4:
            raise ValueError("the gait logging metrics expect four feet in sensor_cfg")
        for k in range(4):
            a.feet_ids[k] = int(ids[k])
        gait = None
        try:
            gait = getattr(env.reward_manager.get_term_cfg("gait").func, "valid_last_air_time", None)  # commands.py:399-401
        except (KeyError, AttributeError):
            pass
        a.gait_valid_last_air_time = _C.ptr(gait, torch.float32, "valid_last_air_time") if gait is not None else None
        a.u = _C.ptr(u, torch.float32, "u") if u is not None else None
        a.seed, a.offset, a.offset_base = self.seed, self.calls, None
        a.workspace, a.workspace_bytes = _C.ptr(self._ws), self._ws.numel()
        return a

    def _set_mask(self, env_ids):
        if env_ids is None:
            self._mask.fill_(1)
        elif torch.is_tensor(env_ids) and env_ids.dtype in (torch.bool, torch.uint8):
            self._mask.copy_(env_ids.view(torch.uint8) if env_ids.dtype == torch.bool else env_ids)
        else:
            ids = torch.as_tensor(env_ids, device=self.device, dtype=torch.long)
            self._mask.zero_()
            self._mask[ids] = 1

    def reset(self, env_ids=None, u: torch.Tensor | None = None, want_extras: bool = True) -> dict:
        """[IL] CommandTerm.reset.  ``want_extras=False`` skips the one host read (the logged means of the metric rows)."""
        self._set_mask(env_ids)
        self._extras.zero_()
        a = self._args(_C.LT_CMD_RESET, 0.0, u)
        check(lib().lt_command_step(C.byref(a), current_stream()), "lt_command_step[reset]")
        count_launches()
        self.calls += 1
        if not want_extras:
            return {}
        ex = self._extras.cpu()
        cnt = float(ex[-1])
        return {k: (float(ex[i]) / cnt if cnt > 0 else float("nan")) for i, k in enumerate(_C.LT_CMD_METRICS)}

    def compute(self, dt: float, u: torch.Tensor | None = None):
        """[IL] CommandTerm.compute: metrics, timer, resampling, zero-command / standing rules -- one launch (+ the row broadcast)."""
        a = self._args(_C.LT_CMD_COMPUTE, dt, u)
        check(lib().lt_command_step(C.byref(a), current_stream()), "lt_command_step[compute]")
        count_launches(2)
        self.calls += 1


class _CfgView:
    """The reference mutates ``cfg.ranges`` / ``cfg.previous_ranges`` / ``cfg.rel_standing_envs`` in place and other code
    reads them back (curriculums.py:189, play scripts): same names, backed by the device block; everything else passes through."""

    def __init__(self, term, cfg):
        object.__setattr__(self, "_term", term)
        object.__setattr__(self, "_cfg", cfg)
        object.__setattr__(self, "ranges", _RangesView(term, "ranges"))
        object.__setattr__(self, "previous_ranges", _RangesView(term, "previous"))

    @property
    def rel_standing_envs(self) -> float:
        return float(self._term.read_state().rel_standing_envs)

    def __getattr__(self, key):
        return getattr(self._cfg, key)

    def __setattr__(self, key, value):
        if key == "rel_standing_envs":
            st = self._term.read_state()
            st.rel_standing_envs = float(value)
            self._term.write_state(st)
        else:
            setattr(self._cfg, key, value)
