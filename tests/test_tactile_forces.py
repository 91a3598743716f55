"""Force-valued tactile encodings: Processed / Normalized / Discrete / Continuous tactile signals against
the oracle and the vectors of the unmodified reference classes (observations.py:311-429)."""
import os

import numpy as np
import pytest
import torch

from oracle import tactile as OT
from tests import helpers as H
from tests.golden.make_golden import TACTILE_FORCE, TACTILE_FORCE_PARAMS, tactile_force_uniforms

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "tactile_force_c4.npz")
KW = dict(p_drop=0.15, p_add=0.05, add_force_noise=True, force_n_prop_min=-0.6, force_n_prop_max=0.1, maximal_force=1.5, total_levels=5,
          add_level_noise=True, level_n_min=-1, level_n_max=1)


def _inputs():
    from locotouch_b200.sim import synth

    c = TACTILE_FORCE
    env = synth.make_env(c["N"], seed=c["seed"], with_object=True, with_tactile=True, tactile_jitter=0.2)
    for step in range(c["steps"]):
        quat = env.scene["robot"].data.body_quat_w[:, synth.NUM_ROBOT_BODIES:].clone()
        force = env.scene.sensors["tactile_contact_sensor"].data.net_forces_w.clone()
        yield step, env, quat, force, tactile_force_uniforms(step, c["N"])
        synth.advance(env, tactile_jitter=0.2)


def test_oracle_matches_reference_golden():
    g = np.load(GOLDEN)
    thr = torch.from_numpy(g["thresholds"])
    N = TACTILE_FORCE["N"]
    for step, _, quat, force, u in _inputs():
        o = OT.force_signals(quat, force, thr, u, **KW)
        ref = torch.from_numpy(g[f"processed_{step}"]).view(N, 4, 221)
        H.assert_equal(o["contact"].float(), ref[:, 0], f"step {step}: processed contact taxels")
        H.assert_close(o["normalized"], ref[:, 1], f"step {step}: normalised forces", rtol=1e-6, atol=1e-7)
        H.assert_close(o["minmax"], ref[:, 2], f"step {step}: min-max normalised signals", rtol=1e-6, atol=1e-7)
        H.assert_close(o["discretized"], ref[:, 3], f"step {step}: discretised signals", rtol=1e-6, atol=1e-7)
        if step == 0:
            for key, ch in (("normalized_0", "minmax"), ("discrete_0", "discretized"), ("continuous_0", "normalized")):
                two = torch.from_numpy(g[key]).view(N, 2, 221)
                H.assert_equal(two[:, 0], o["contact"].float(), key + " channel 0")
                H.assert_close(two[:, 1], o[ch], key + " channel 1", rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_force_encodings_match_reference_golden(cuda, lt_lib):
    from types import SimpleNamespace

    from locotouch_b200.mdp import observations as O
    from locotouch_b200.sim.scene import SceneEntityCfg

    g = np.load(GOLDEN)
    N = TACTILE_FORCE["N"]
    thr = torch.from_numpy(g["thresholds"]).to(cuda)
    terms = None
    for step, env, quat, force, u in _inputs():
        denv = env.to(cuda)
        params = dict(asset_cfg=SceneEntityCfg("robot", body_names="sensor_.*").resolve(denv.scene),
                      sensor_cfg=SceneEntityCfg("tactile_contact_sensor", body_names="sensor_.*").resolve(denv.scene), **TACTILE_FORCE_PARAMS)
        if terms is None:
            terms = {name: getattr(O, name)(SimpleNamespace(params=params), denv) for name in
                     ("ProcessedTactileSignals", "NormalizedTactileSignals", "DiscreteTactileSignals", "CotinuousTactileSignals")}
            for t in terms.values():
                assert float(t.contact_threshold_envs_sensors.min()) >= 0.04 - 1e-6 and float(t.contact_threshold_envs_sensors.max()) <= 0.06 + 1e-6
                t.contact_threshold_envs_sensors = thr.view(N, 17, 13).clone()
        else:
            for t in terms.values():  # the terms keep the scene handles of the env they were built on: move the new state in
                t.asset.data.body_quat_w.copy_(denv.scene["robot"].data.body_quat_w)
                t.contact_sensor.data.net_forces_w.copy_(denv.scene.sensors["tactile_contact_sensor"].data.net_forces_w)
        ud = {k: v.to(cuda) for k, v in u.items()}
        sig = terms["ProcessedTactileSignals"](denv, u=ud, **params)
        ref = torch.from_numpy(g[f"processed_{step}"])
        assert sig.shape == ref.shape == (N, 884)
        H.assert_equal(sig.view(N, 4, 221)[:, 0].cpu(), ref.view(N, 4, 221)[:, 0], f"step {step}: contact taxels")
        H.assert_close(sig.cpu(), ref, f"step {step}: processed tactile signals", rtol=1e-6, atol=1e-7)
        if step == 0:
            for name, key in (("NormalizedTactileSignals", "normalized_0"), ("DiscreteTactileSignals", "discrete_0"), ("CotinuousTactileSignals", "continuous_0")):
                out = terms[name](denv, u=ud, **params)
                H.assert_close(out.cpu(), torch.from_numpy(g[key]), key, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 405, 4097])
def test_force_encodings_match_oracle_on_ragged_sizes_and_rng_path(cuda, lt_lib, n):
    from locotouch_b200 import ops
    from locotouch_b200.sim import synth

    env = synth.make_env(n, seed=300 + n, with_object=True, with_tactile=True, tactile_jitter=0.3)
    quat_all = env.scene["robot"].data.body_quat_w
    force = env.scene.sensors["tactile_contact_sensor"].data.net_forces_w
    g = torch.Generator().manual_seed(n)
    thr = 0.05 + (torch.rand(n, 221, generator=g) * 0.02 - 0.01)
    u = {k: torch.rand(n, 221, generator=g) for k in ("drop", "drop_force", "add", "add_force", "noise", "small", "level")}
    o = OT.force_signals(quat_all[:, synth.NUM_ROBOT_BODIES:], force, thr, u, **KW)
    out = torch.zeros(n, 4, 221, device=cuda)
    kw = dict(p_drop=0.15, p_add=0.05, add_force_noise=True, force_n_prop_min=-0.6, force_n_prop_max=0.1, maximal_force=1.5, total_levels=5,
              add_level_noise=True, level_n_min=-1, level_n_max=1)
    ops.taxel_forces(quat_all.to(cuda), force.to(cuda), thr.to(cuda), out, ("contact", "normalized", "minmax", "discretized"),
                     quat_body_offset=synth.NUM_ROBOT_BODIES, u={k: v.to(cuda) for k, v in u.items()}, **kw)
    H.assert_equal(out[:, 0].cpu(), o["contact"].float(), "contact")
    H.assert_close(out[:, 1].cpu(), o["normalized"], "normalised", rtol=1e-6, atol=1e-7)
    H.assert_close(out[:, 2].cpu(), o["minmax"], "min-max", rtol=1e-6, atol=1e-7)
    H.assert_close(out[:, 3].cpu(), o["discretized"], "discretised", rtol=1e-6, atol=1e-7)
    # in-kernel Philox: deterministic per (seed, offset), different across offsets, rates near the configured probabilities
    a, b, c2 = (torch.zeros(n, 4, 221, device=cuda) for _ in range(3))
    for buf, off in ((a, 3), (b, 3), (c2, 4)):
        ops.taxel_forces(quat_all.to(cuda), force.to(cuda), thr.to(cuda), buf, ("contact", "normalized", "minmax", "discretized"),
                         quat_body_offset=synth.NUM_ROBOT_BODIES, seed=7, offset=off, **kw)
    H.assert_equal(a, b, "same (seed, offset) -> same draw")
    if n >= 405:
        assert not torch.equal(a, c2)
        orig = o["original_contact"].float().mean().item()
        assert abs(a[:, 0].mean().item() - (orig * 0.85 + (1 - orig * 0.85) * 0.05)) < 0.01
    assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0
