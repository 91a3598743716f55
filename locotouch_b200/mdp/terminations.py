"""Termination terms with the reference's names / signatures (reference locotouch/mdp/terminations.py:10-23); bool rows of
the fused launch (K1)."""
from __future__ import annotations

import torch

from . import task_spec as TS
from ._fusion import termination_term

__all__ = ["object_below_robot", "bad_roll"]


def object_below_robot(env, robot_cfg=None, object_cfg=None) -> torch.Tensor:
    return termination_term(env, TS.TK_OBJECT_BELOW_ROBOT)


def bad_roll(env, limit_angle: float, asset_cfg=None) -> torch.Tensor:
    """asin(projected_gravity_b.y).abs() > limit_angle on the transported object (the only asset the LocoTouch cfgs use it
    with, reference cylinder_transport_teacher_env_cfg.py:47-52)."""
    if asset_cfg is not None and getattr(asset_cfg, "name", "object") != "object":
        raise ValueError("the fused bad_roll term is evaluated on the 'object' asset")
    return termination_term(env, TS.TK_BAD_ROLL, (limit_angle,))
