"""ReplayBuffer with the reference's interface (reference locotouch/distill/replay_buffer.py:9-150), device resident.

Same constructor, ``collect_data(teacher_policy, student_policy, num_steps) -> (trajectory_rewards, trajectory_lengths)``,
``to_recurrent_generator(batch_size)`` yielding ``dict(proprioceptions [L, B, P], teacher_encoder_obses [L, B, E],
tactile_signals [L, B, ...], masks [L, B] bool)``, ``clear_buffer``, ``evaluate``, ``num_trajs``, ``num_steps``.  Below it:

* per env step the observations are appended to step-major device buffers ``[S, N, D]`` (three row copies) and ONE kernel
  (K11 ``lt_dagger_step``) does the done handling -- reward sums, episode log, the "record while the step budget lasts, in
  env order" rule -- where the reference runs ``dones.any()``, ``nonzero``, two ``.cpu().tolist()`` and a Python loop with
  ``.item()`` and per-trajectory ``torch.stack`` of per-step slices; the host reads back four integers per step (the loop
  condition needs the recorded-step count);
* finished trajectories are packed once per collection into flat ``[total_steps, D]`` stores (K11 ``lt_pack_trajectories``);
* every batch is padded by the K8 kernel (``lt_pad_trajectories``) instead of a Python loop of slice assignments.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import ops
from .tactile_recorder import TactileRecorder


class _Grow:
    """Append-only device array that doubles its capacity (old contents copied on growth)."""

    def __init__(self, shape_tail, dtype, device, capacity=64):
        self.buf = torch.empty((capacity, *shape_tail), dtype=dtype, device=device)
        self.n = 0

    def reserve(self, extra: int):
        need = self.n + extra
        if need > self.buf.shape[0]:
            cap = max(need, 2 * self.buf.shape[0])
            new = torch.empty((cap, *self.buf.shape[1:]), dtype=self.buf.dtype, device=self.buf.device)
            new[:self.n].copy_(self.buf[:self.n])
            self.buf = new

    def view(self):
        return self.buf[:self.n]


class ReplayBuffer:
    def __init__(self, env, tactile_recorder: TactileRecorder, proprioception_dim: int):
        self._env = env
        self._num_envs = env.num_envs
        self._device = env.device
        self._proprioception_dim = proprioception_dim
        self._tactile_recorder = tactile_recorder
        self._steps_count = 0
        self._reward_sums = torch.zeros(self._num_envs, device=self._device)
        self._flat = None        # (proprioceptions, teacher_encoder_obses, tactile_signals) flat [total_steps, D] stores
        self._tail_shapes = None
        self._lengths: list[int] = []   # host copy (the reference keeps per-trajectory tensors, i.e. their lengths, on the host too)
        self._offsets_dev = None
        self._lengths_dev = None

    # --------------------------------------------------------------------------------------------------- collection
    def _collect_loop(self, observe_and_act, advance, num_steps: Optional[int], num_trajs: Optional[int], record: bool):
        """Shared by collect_data (record=True: budget in steps) and evaluate (record=False: budget in episodes)."""
        N, dev = self._num_envs, self._device
        start_idx = torch.zeros(N, device=dev, dtype=torch.int32)
        state = torch.zeros(4, device=dev, dtype=torch.int64)
        state[0] = self._steps_count
        limit = self._steps_count + num_steps if record else 0
        traj = [_Grow((), torch.int32, dev, 4 * N) for _ in range(3)]  # env, start, length
        eps_r, eps_l = _Grow((), torch.float32, dev, 4 * N), _Grow((), torch.int32, dev, 4 * N)
        steps = None  # step-major observation buffers, allocated at the first step (shapes come from the env)
        steps_count, n_traj, n_eps = 0, 0, 0
        while (self._steps_count < limit) if record else (n_eps < num_trajs):
            rows, action = observe_and_act()
            if record:  # stored before the env moves on (reference replay_buffer.py:44 "store the data before ... is updated")
                if steps is None:
                    steps = [_Grow((N, *r.shape[1:]), torch.float32, dev, 64) for r in rows]
                for g, r in zip(steps, rows):
                    g.reserve(1)
                    g.buf[g.n].copy_(r)
                    g.n += 1
            reward, dones = advance(action)
            if dones.dtype not in (torch.bool, torch.uint8):
                dones = dones != 0
            steps_count += 1
            for g in (*traj, eps_r, eps_l):
                g.reserve(N)
            ops.dagger_step(dones, reward.float(), self._reward_sums, start_idx, steps_count, limit, state, traj[0].buf, traj[1].buf, traj[2].buf,
                            eps_r.buf, eps_l.buf, always_restart=not record)
            self._steps_count, n_traj, n_eps, _ = state.tolist()  # the one device -> host read of the step
            for g in traj:
                g.n = n_traj
            eps_r.n = eps_l.n = n_eps
            yield dones
        self._last = dict(traj=traj, steps=steps, eps_r=eps_r, eps_l=eps_l)

    def collect_data(self, teacher_policy, student_policy: Optional[torch.nn.Module], num_steps: int):
        if student_policy is not None:  # reference replay_buffer.py:22-23
            self._env.reset()
        self._tactile_recorder.reset()
        with torch.no_grad():
            env_obs = self._env.get_observations()
            carry = {"obs": env_obs["policy"], "tactile": env_obs["tactile"]}

            def observe_and_act():
                pos = carry["obs"]
                proprioception = pos[:, :self._proprioception_dim]
                teacher_encoder_obs = pos[:, self._proprioception_dim:]
                tactile_signal = carry["tactile"]
                action = teacher_policy(pos) if student_policy is None else student_policy(proprioception, tactile_signal)
                self._tactile_recorder.record_new_tactile_signals(tactile_signal)
                return (proprioception, teacher_encoder_obs, self._tactile_recorder.get_tactile_signals()), action

            def advance(action):
                next_obs, reward, dones, _ = self._env.step(action)
                carry["obs"], carry["tactile"] = next_obs["policy"], next_obs["tactile"]
                return reward, dones

            for dones in self._collect_loop(observe_and_act, advance, num_steps, None, record=True):
                if student_policy is not None:
                    student_policy.reset(dones)
                self._tactile_recorder.reset_where(dones)
        last = self._last
        self._append_trajectories(last["traj"], last["steps"])
        return last["eps_r"].view().tolist(), last["eps_l"].view().tolist()

    def _append_trajectories(self, traj, steps):
        env, start, length = (g.view() for g in traj)
        M = env.numel()
        if M == 0 or steps is None:
            return
        lengths = length.tolist()
        total = sum(lengths)
        offsets_local = torch.cumsum(length.long(), dim=0) - length.long()
        new_flat = []
        for i, g in enumerate(steps):
            x = g.view()  # [S, N, ...]
            x2 = x.reshape(x.shape[0], x.shape[1], -1)
            old = self._flat[i] if self._flat is not None else None
            rows_before = old.shape[0] if old is not None else 0
            flat = torch.empty(rows_before + total, x2.shape[2], device=self._device)
            if old is not None:
                flat[:rows_before].copy_(old)
            ops.pack_trajectories(x2, env, start, offsets_local, total, flat[rows_before:])
            new_flat.append(flat)
        self._flat = new_flat
        self._tail_shapes = [tuple(g.buf.shape[2:]) for g in steps]
        self._lengths.extend(lengths)
        ln = torch.tensor(self._lengths, dtype=torch.int64, device=self._device)
        self._lengths_dev = ln
        self._offsets_dev = torch.cumsum(ln, dim=0) - ln

    # ------------------------------------------------------------------------------------------------------ batches
    def to_recurrent_generator(self, batch_size: int):
        num_trajs = len(self._lengths)
        traj_indices = np.random.permutation(np.arange(num_trajs))  # the reference's draw (replay_buffer.py:84-85)
        for start_idx in range(0, num_trajs, batch_size):
            end_idx = np.minimum(start_idx + batch_size, num_trajs)
            yield self._prepare_padded_sequence(traj_indices[start_idx:end_idx])

    def _prepare_padded_sequence(self, traj_indices):
        max_length = max(self._lengths[int(i)] for i in traj_indices)
        idx = torch.as_tensor(np.asarray(traj_indices), dtype=torch.int64, device=self._device)
        offsets, lengths = self._offsets_dev[idx].contiguous(), self._lengths_dev[idx].contiguous()
        outs, masks = [], None
        for flat, tail in zip(self._flat, self._tail_shapes):
            out, masks = ops.pad_trajectories(flat, offsets, lengths, max_length)
            outs.append(out.view(max_length, len(traj_indices), *tail))
        return dict(proprioceptions=outs[0], teacher_encoder_obses=outs[1], tactile_signals=outs[2], masks=masks)

    def clear_buffer(self):
        self._flat, self._lengths, self._offsets_dev, self._lengths_dev = None, [], None, None
        self._steps_count = 0
        self._reward_sums[:] = 0

    # --------------------------------------------------------------------------------------------------- evaluation
    def evaluate(self, student_policy, num_trajs: int):
        """reference replay_buffer.py:118-140 (tuple-style ``get_observations`` / ``step`` of the RSL-RL wrapper)."""
        with torch.no_grad():
            obs, extras = self._env.get_observations()
            carry = {"obs": obs, "tactile": extras["observations"]["tactile"]}

            def observe_and_act():
                pos = carry["obs"]
                return (), student_policy(pos[:, :self._proprioception_dim], carry["tactile"])

            def advance(action):
                next_obs, reward, dones, extras = self._env.step(action)
                carry["obs"], carry["tactile"] = next_obs, extras["observations"]["tactile"]
                return reward, dones

            saved = self._steps_count
            for _ in self._collect_loop(observe_and_act, advance, None, num_trajs, record=False):
                pass
            self._steps_count = saved
        last = self._last
        return last["eps_r"].view().tolist(), [float(v) for v in last["eps_l"].view().tolist()]

    @property
    def num_trajs(self):
        return len(self._lengths)

    @property
    def num_steps(self):
        return self._steps_count
