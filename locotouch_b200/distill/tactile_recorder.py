"""TactileRecorder with the reference's interface (reference locotouch/distill/tactile_recorder.py:4-34) on K2's generic
fp32 delay-line kernel: one launch per ``record_new_tactile_signals`` instead of a full-buffer clone + nonzero + scatter."""
from __future__ import annotations

import torch

from .. import ops


class TactileRecorder:
    def __init__(self, device, env_num, tactile_shape, min_delay=3, max_delay=7):
        self.device = device
        self.env_num = env_num
        self.tactile_shape = (tactile_shape,) if isinstance(tactile_shape, int) else tuple(tactile_shape)
        self.min_delay = min_delay
        self.max_delay = max_delay
        self._dim = 1
        for s in self.tactile_shape:
            self._dim *= s
        self.tactile_buffer = torch.zeros((env_num, max_delay, *self.tactile_shape), dtype=torch.float32, device=device)
        self.first_signal_recorded = torch.ones((env_num,), dtype=torch.bool, device=device)
        self.delay_steps = torch.zeros((env_num,), dtype=torch.long, device=device)
        self.env_idx = torch.arange(env_num, device=device)
        self._out = torch.zeros((env_num, *self.tactile_shape), dtype=torch.float32, device=device)
        self.reset()

    def reset(self, env_idx=None):
        env_idx = env_idx if env_idx is not None else self.env_idx
        self.tactile_buffer[env_idx] = 0.0
        self.first_signal_recorded[env_idx] = True
        # exclusive high, exactly like the reference (tactile_recorder.py:22): delay in [min_delay, max_delay - 1]
        self.delay_steps[env_idx] = torch.randint(low=self.min_delay, high=self.max_delay, size=(env_idx.shape), device=self.device)

    def reset_where(self, mask: torch.Tensor):
        """``reset(mask.nonzero())`` without the host synchronisation of ``nonzero``: one delay is drawn per env and kept only
        where ``mask`` is set (same distribution; the reference draws one per reset env)."""
        m = mask.view(-1).bool()
        self.tactile_buffer.masked_fill_(m.view(-1, *([1] * (self.tactile_buffer.dim() - 1))), 0.0)
        self.first_signal_recorded.logical_or_(m)
        draw = torch.randint(low=self.min_delay, high=self.max_delay, size=(self.env_num,), device=self.device)
        self.delay_steps.copy_(torch.where(m, draw, self.delay_steps))

    def record_new_tactile_signals(self, tactile_signals: torch.Tensor):
        ops.tactile_delay(self.tactile_buffer.view(self.env_num, self.max_delay, self._dim), self.first_signal_recorded, self.delay_steps,
                          tactile_signals.contiguous().view(self.env_num, self._dim), self._out.view(self.env_num, self._dim))

    def get_tactile_signals(self):
        return self._out.clone()
