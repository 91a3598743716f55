"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement of the binary tactile path:

* ``normal_forces`` / ``binary_taxels``  reference locotouch/mdp/observations.py:154-159 (get_original_signals),
  :166-199 (get_normal_forces), :281-308 (BinaryTactileSignals.__call__)
* ``TactileDelayOracle``                 reference locotouch/distill/tactile_recorder.py:4-34

The reference draws its dropout / addition masks with ``torch.rand_like`` (observations.py:173,181); a bit-exact
bitmap is only defined when those uniforms are an explicit input, so they are arguments here (``u_drop``,
``u_add``).  tests/test_oracle_vs_reference.py feeds the live reference the same uniforms by patching
``torch.rand_like`` and checks equality bit for bit.
"""
from __future__ import annotations

import torch

from . import il_math as M


def normal_forces(body_quat_w, net_forces_w):
    """F_n = -(R(q)^T F_w).z per taxel.  [N,T,4] x [N,T,3] -> [N,T]   (observations.py:156-158)."""
    return -M.quat_apply_inverse(body_quat_w, net_forces_w)[..., 2]


def binary_taxels(body_quat_w, net_forces_w, thresholds, u_drop=None, u_add=None, p_drop=0.005, p_add=0.005):
    """-> (contact bool [N,T], signal f32 [N,2*T]).  Strict '>' at the threshold (observations.py:159)."""
    fn = normal_forces(body_quat_w, net_forces_w)
    original = fn > thresholds
    contact = original.clone()
    if p_drop > 0.0 and u_drop is not None:  # observations.py:172-176
        contact = contact & ~(u_drop < p_drop)
    if p_add > 0.0 and u_add is not None:  # observations.py:180-185 (a just-dropped taxel can be re-added)
        contact = contact | (u_add < p_add)
    sig = contact.float()
    return dict(original=original, contact=contact, normal_forces=fn, signal=torch.cat([sig, sig], dim=1))


def pack_bits(contact):
    """[N,T] bool -> [N, ceil(T/32)] int32 little-endian bit order (taxel t -> word t//32, bit t%32)."""
    n, t = contact.shape
    words = (t + 31) // 32
    pad = torch.zeros(n, words * 32, dtype=torch.int64)
    pad[:, :t] = contact.long()
    w = (pad.view(n, words, 32) << torch.arange(32).view(1, 1, 32)).sum(dim=-1)
    w = torch.where(w >= 2**31, w - 2**32, w)
    return w.to(torch.int32)


class TactileDelayOracle:
    """tactile_recorder.py:4-34 with the per-env delay as explicit state (``randint(min, max)`` with exclusive high)."""

    def __init__(self, num_envs, dim, min_delay=1, max_delay=2, delay_steps=None):
        self.buf = torch.zeros(num_envs, max_delay, dim)
        self.first = torch.ones(num_envs, dtype=torch.bool)
        self.max_delay = max_delay
        self.delay = torch.full((num_envs,), min_delay, dtype=torch.long) if delay_steps is None else delay_steps.clone()

    def reset(self, ids, delay_steps=None):
        self.buf[ids] = 0.0
        self.first[ids] = True
        if delay_steps is not None:
            self.delay[ids] = delay_steps

    def record(self, x):
        self.buf[:, 1:] = self.buf[:, :-1].clone()
        self.buf[:, 0] = x
        f = self.first
        self.buf[f] = x[f].unsqueeze(1).expand(-1, self.max_delay, -1)
        self.first[:] = False

    def get(self):
        return self.buf[torch.arange(self.buf.shape[0]), self.delay]


def force_signals(body_quat_w, net_forces_w, thresholds, u, p_drop=0.0, p_add=0.0, add_force_noise=False, force_n_prop_min=0.0,
                  force_n_prop_max=0.0, maximal_force=1.0, total_levels=5, add_level_noise=False, level_n_min=0.0, level_n_max=0.0):
    """Force-valued tactile encodings: reference observations.py:166-199 (get_normal_forces: dropout / addition with synthetic
    forces, proportional force noise), :201-205 (normalised forces), :207-224 (per-env min-max normalisation), :226-237
    (discretisation with level noise).  ``u`` = dict of explicit [N, T] uniforms standing for the reference's ``rand_like`` draws:
    drop, drop_force, add, add_force, noise, small, level (the reference draws the masked ones only for the selected taxels, in
    row-major order -- element (n, t) of the full tensor is the draw that taxel would get).  Returns every intermediate the
    reference keeps (processed_* attributes) plus the masks that selected the draws."""
    fn = normal_forces(body_quat_w, net_forces_w)
    contact = fn > thresholds
    out = dict(original_contact=contact.clone(), original_normal_forces=fn.clone())
    drop = torch.zeros_like(contact)
    if p_drop > 0.0:
        drop = contact & (u["drop"] < p_drop)
        fn = torch.where(drop, u["drop_force"] * thresholds, fn)
        contact = contact & ~drop
    add = torch.zeros_like(contact)
    if p_add > 0.0:
        add = ~contact & (u["add"] < p_add)
        fn = torch.where(add, thresholds * (1.0 + 0.2 * u["add_force"]), fn)
        contact = contact | add
    small = torch.zeros_like(contact)
    if add_force_noise:
        fn = torch.where(contact, fn * (1.0 + (u["noise"] * (force_n_prop_max - force_n_prop_min) + force_n_prop_min)), fn)
        fn = torch.clamp(fn, min=0.0)
        small = contact & (fn < thresholds)
        fn = torch.where(small, thresholds * (1.0 + 0.2 * u["small"]), fn)
    normalized = torch.clamp(fn / maximal_force, 0.0, 1.0)
    valid = torch.where(contact, normalized, torch.zeros_like(normalized))
    mn = valid.min(dim=-1, keepdim=True)[0]
    mx = valid.max(dim=-1, keepdim=True)[0]
    rng = torch.where((mx - mn) > 0.0, mx - mn, torch.ones_like(mx))
    minmax = torch.clamp((valid - mn) / rng, 0.0, 1.0)
    bin_ = 1.0 / total_levels
    disc = torch.round(minmax / bin_)
    if add_level_noise:
        disc = disc + (u["level"] * (level_n_max - level_n_min) + level_n_min)
    disc = torch.clamp(disc * bin_, 0.0, 1.0)
    disc = torch.where(contact, disc, torch.zeros_like(disc))
    out.update(contact=contact, normal_forces=fn, normalized=normalized, minmax=minmax, discretized=disc, drop_mask=drop, add_mask=add,
               noise_mask=contact.clone(), small_mask=small)
    return out
