"""Per-term drop-in plumbing: one fused launch per env step behind IsaacLab's one-term-at-a-time manager API
(SURVEY.md 8b "Fusion under a per-term API")."""
from __future__ import annotations

import copy

import torch

from . import task_spec as TS
from .fused import FusedMdp

_ATTR = "_locotouch_b200_fused"


def _has_object(env) -> bool:
    try:
        env.scene["object"]
        return True
    except (KeyError, AttributeError):
        return False


def per_term_spec(env) -> TS.TaskSpec:
    """All reward kinds the scene supports, weight 1 (the manager applies the real weight), default task parameters."""
    spec = getattr(env, "lt_task_spec", None)
    if spec is None:
        spec = TS.teacher_spec() if _has_object(env) else TS.locomotion_spec()
    spec = copy.deepcopy(spec)
    if spec.with_object:
        have = {t.kind for t in spec.rewards}
        extra = [("object_rp_angle", TS.RK_OBJ_RP_ANGLE), ("object_rp_velocity", TS.RK_OBJ_RP_VEL)]
        spec.rewards += [TS.RewardTerm(n, k, 1.0) for n, k in extra if k not in have]
    for t in spec.rewards:
        t.weight = 1.0
    spec.max_episode_length = int(getattr(env, "max_episode_length", spec.max_episode_length))
    return spec


class _Cache:
    def __init__(self, env):
        self.fused = FusedMdp(env, per_term_spec(env))
        self.step = None
        self.obj_obs = None
        self.obj_step = None


def cache_for(env) -> _Cache:
    c = getattr(env, _ATTR, None)
    if c is None:
        c = _Cache(env)
        setattr(env, _ATTR, c)
    return c


def reward_term(env, kind: int, params: tuple = ()) -> torch.Tensor:
    """Row of the fused result for ``kind``; launches the fused reward pass on the first term call of an env step."""
    c = cache_for(env)
    f = c.fused
    idx = next((i for i, t in enumerate(f.spec.rewards) if t.kind == kind), None)
    if idx is None:
        raise ValueError(f"reward kind {kind} is not part of the fused table of this scene")
    want = tuple(float(x) for x in params)
    have = tuple(float(x) for x in f.spec.rewards[idx].p[: len(want)])
    if want and any(abs(a - b) > 1e-12 for a, b in zip(want, have)):
        # parameters differ from the fused table: update the table (one-off) and recompute this step
        p = list(f.spec.rewards[idx].p) + [0.0] * (len(want) - len(f.spec.rewards[idx].p))
        p[: len(want)] = want
        f.spec.rewards[idx].p = tuple(p)
        for k, v in enumerate(p):
            f._args.reward_terms[idx].p[k] = v
        c.step = None
    step = getattr(env, "common_step_counter", None)
    if c.step is None or step is None or step != c.step:
        f.env = env
        f.compute_rewards(auto_reset=False)
        c.step = step
    return f.term_raw[idx]


def termination_term(env, kind: int, params: tuple = ()) -> torch.Tensor:
    c = cache_for(env)
    f = c.fused
    idx = next((i for i, t in enumerate(f.spec.terminations) if t.kind == kind), None)
    if idx is None:
        raise ValueError(f"termination kind {kind} is not part of the fused table of this scene")
    if params and abs(float(f.spec.terminations[idx].p[0]) - float(params[0])) > 1e-12:
        f.spec.terminations[idx].p = (float(params[0]),)
        f._args.termination_terms[idx].p[0] = float(params[0])
        c.step = None
    step = getattr(env, "common_step_counter", None)
    if c.step is None or step is None or step != c.step:
        f.env = env
        f.compute_rewards(auto_reset=False)
        c.step = step
    return f.term_masks[idx]


def reset_terms(env, env_ids=None):
    cache_for(env).fused.reset(env_ids)
