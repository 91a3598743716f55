"""ContactSensor state with IsaacLab's attribute names, advanced by ONE kernel per step (K18, ``lt_contact_sensor_update``).

Stands where ``env.scene.sensors[name]`` stands for the reference terms (reference locotouch/mdp/rewards.py:116-156,596-604,
observations.py:60-66 read ``.data.current_air_time / current_contact_time / last_air_time / last_contact_time /
net_forces_w / net_forces_w_history``): [IL] ``ContactSensor._update_buffers_impl`` + ``reset`` without the
~10 boolean-index / where / clone launches per sensor and step.  PhysX still produces the net contact forces; ``update`` takes them."""
from __future__ import annotations

from types import SimpleNamespace

import torch

from .. import ops


class ContactSensorState:
    def __init__(self, num_envs: int, body_names, history_length: int = 3, force_threshold: float = 1.0, device="cuda"):
        self.body_names = list(body_names)
        self.num_bodies = len(self.body_names)
        self.cfg = SimpleNamespace(history_length=history_length, force_threshold=force_threshold, track_air_time=True)
        z = lambda *shape: torch.zeros(*shape, device=device)  # noqa: E731
        n, b = num_envs, self.num_bodies
        self.data = SimpleNamespace(net_forces_w=z(n, b, 3), net_forces_w_history=z(n, max(history_length, 1), b, 3), current_air_time=z(n, b),
                                    last_air_time=z(n, b), current_contact_time=z(n, b), last_contact_time=z(n, b))

    def find_bodies(self, pattern):
        import re

        pats = [pattern] if isinstance(pattern, str) else list(pattern)
        ids = [i for i, nm in enumerate(self.body_names) if any(re.fullmatch(p, nm) for p in pats)]
        return ids, [self.body_names[i] for i in ids]

    def update(self, forces: torch.Tensor, dt: float, reset_mask: torch.Tensor | None = None):
        """One sensor update with this step's net contact forces [N, bodies, 3]; envs flagged in ``reset_mask`` (uint8 [N]) are
        cleared instead (``ContactSensor.reset(env_ids)`` folded into the same launch)."""
        d = self.data
        ops.contact_sensor_update(forces.contiguous(), net_forces_w=d.net_forces_w, history=d.net_forces_w_history if self.cfg.history_length > 0 else None,
                                  current_air_time=d.current_air_time, last_air_time=d.last_air_time, current_contact_time=d.current_contact_time,
                                  last_contact_time=d.last_contact_time, dt=dt, force_threshold=self.cfg.force_threshold, reset_mask=reset_mask)

    def reset(self, env_ids=None):
        d = self.data
        idx = slice(None) if env_ids is None else env_ids
        for t in (d.net_forces_w, d.net_forces_w_history, d.current_air_time, d.last_air_time, d.current_contact_time, d.last_contact_time):
            t[idx] = 0.0

    def compute_first_contact(self, dt: float, abs_tol: float = 1.0e-8):
        """[IL] ContactSensor.compute_first_contact: bodies that came into contact within the last ``dt`` seconds."""
        c = self.data.current_contact_time
        return (c > 0.0) & (c < dt + abs_tol)

    def compute_first_air(self, dt: float, abs_tol: float = 1.0e-8):
        a = self.data.current_air_time
        return (a > 0.0) & (a < dt + abs_tol)
