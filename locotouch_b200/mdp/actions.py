"""JointPositionActionPrevPrev state machine (reference locotouch/mdp/actions.py:13-69) on the K0 kernel.

The reference class derives from IsaacLab's JointPositionAction; this mirror keeps its buffers / properties
(``raw_actions``, ``prev_raw_actions``, ``prev_prev_raw_actions``, ``processed_actions`` ...) and ``process_actions`` /
``reset`` semantics without the simulator: ``apply_actions`` (writing joint targets into PhysX) stays IsaacLab's.
"""
from __future__ import annotations

import torch

from .. import ops


class JointPositionActionPrevPrev:
    def __init__(self, num_envs: int, action_dim: int, device, scale: float = 1.0, offset: torch.Tensor | float = 0.0,
                 clip_raw_actions: bool = False, raw_action_clip_value: float = 100.0, raw_action_scale: float = 1.0):
        z = lambda: torch.zeros(num_envs, action_dim, device=device)  # noqa: E731
        self._raw_actions, self._prev_raw_actions, self._prev_prev_raw_actions = z(), z(), z()
        self._scale = scale
        self._offset = offset if torch.is_tensor(offset) else torch.full((num_envs, action_dim), float(offset), device=device)
        self._clip_raw_actions = clip_raw_actions
        self._raw_action_clip_value = raw_action_clip_value
        self._raw_action_scale = raw_action_scale
        self._processed_actions = self._raw_actions * self._scale + self._offset
        self._prev_processed_actions = self._processed_actions.clone()
        self._prev_prev_processed_actions = self._processed_actions.clone()

    def process_actions(self, actions: torch.Tensor):
        ops.process_actions(actions.contiguous(), self._raw_actions, self._prev_raw_actions, self._prev_prev_raw_actions,
                            self._processed_actions, self._prev_processed_actions, self._prev_prev_processed_actions,
                            clip=self._raw_action_clip_value if self._clip_raw_actions else 0.0, raw_scale=self._raw_action_scale,
                            scale=self._scale, offset=self._offset)

    def reset(self, env_ids=None):
        ids = slice(None) if env_ids is None else env_ids
        self._prev_raw_actions[ids] = 0.0
        self._prev_prev_raw_actions[ids] = 0.0
        self._processed_actions[ids] = (self._raw_actions * self._scale + self._offset)[ids]
        self._prev_processed_actions[ids] = self._processed_actions[ids].clone()
        self._prev_prev_processed_actions[ids] = self._processed_actions[ids].clone()
        self._raw_actions[ids] = 0.0  # [IL] ActionTerm.reset

    raw_actions = property(lambda self: self._raw_actions)
    processed_actions = property(lambda self: self._processed_actions)
    prev_raw_actions = property(lambda self: self._prev_raw_actions)
    prev_prev_raw_actions = property(lambda self: self._prev_prev_raw_actions)
    prev_processed_actions = property(lambda self: self._prev_processed_actions)
    prev_prev_processed_actions = property(lambda self: self._prev_prev_processed_actions)
