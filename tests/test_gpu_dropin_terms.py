"""Per-term drop-in API (manager-term signatures), TactileRecorder and the student-batch helpers on the GPU."""
import math
from types import SimpleNamespace

import pytest
import torch

from oracle import tactile as OT
from oracle.mdp import MdpOracle
from tests import helpers as H
from locotouch_b200.mdp import task_spec as TS
from locotouch_b200.sim import synth
from locotouch_b200.sim.scene import SceneEntityCfg

pytestmark = pytest.mark.gpu


def test_per_term_calls_return_rows_of_one_fused_launch(cuda, lt_lib):
    from locotouch_b200 import _C
    from locotouch_b200 import mdp as _  # noqa: F401
    from locotouch_b200.mdp import rewards as R
    from locotouch_b200.mdp import terminations as Tm

    env = synth.make_env(300, seed=8, with_object=True)
    spec = TS.teacher_spec()
    oracle = MdpOracle(env, spec)
    denv = env.to(cuda)
    cfg = H.gait_cfg(spec)
    gait = R.AdaptiveSymmetricGaitRewardwithObject(cfg, denv)
    with pytest.raises(ValueError):
        R.AdaptiveSymmetricGaitReward(cfg, denv)  # scene has an object: the plain class does not match
    for step in range(3):
        out = oracle.step(env, auto_reset=False)
        before = _C.launch_count
        vals = {
            "track_lin_vel_xy": R.track_lin_vel_xy_pst(denv, sigma=0.25), "track_ang_vel_z": R.track_ang_vel_z_pst(denv, sigma=0.25),
            "foot_slip": R.foot_slipping_ngt(denv, threshold=0.5), "foot_dragging": R.foot_dragging_ngt(denv, height_threshold=0.03, foot_vel_xy_threshold=0.1),
            "gait": gait(denv, **cfg.params), "track_base_height": R.track_base_height_ngt(denv, target_height=0.42),
            "base_z_velocity": R.base_z_velocity_ngt(denv), "base_roll_pitch_angle": R.base_roll_pitch_angle_ngt(denv),
            "base_roll_pitch_velocity": R.base_roll_pitch_velocity_ngt(denv), "joint_position_limit": R.joint_position_limit_ngt(denv),
            "joint_position": R.joint_position_ngt(denv, stand_still_scale=5.0, velocity_threshold=0.3), "joint_acceleration": R.joint_acceleration_ngt(denv),
            "joint_velocity": R.joint_velocity_ngt(denv), "joint_torque": R.joint_torque_ngt(denv), "action_rate": R.action_rate_ngt(denv),
            "thigh_calf_collision": R.thigh_calf_collision_ngt(denv, threshold=0.1),
            "object_xy_position": R.object_relative_xy_position_ngt(denv, work_only_when_cmd=1), "object_z_velocity": R.object_relative_z_velocity_ngt(denv),
            "object_roll_pitch_angle": R.object_relative_roll_angle_ngt(denv), "object_roll_pitch_velocity": R.object_relative_roll_velocity_ngt(denv),
            "object_yaw_alignment": R.object_relative_yaw_angle_ngt(denv, work_only_when_cmd=1),
            "object_dangerous_state": R.object_dangerous_state_ngt(denv, x_max=0.125, y_max=0.097, z_min=0.095, roll_pitch_max=None, vel_xy_max=2.5),
        }
        below, roll = Tm.object_below_robot(denv), Tm.bad_roll(denv, limit_angle=math.pi / 3, asset_cfg=SceneEntityCfg("object"))
        launches = _C.launch_count - before
        assert launches <= 2, f"{launches} launches for 24 term calls: the terms must share one fused launch (+ the any() pre-pass)"
        for name, v in vals.items():
            assert v.shape == (300,)
            H.assert_close(v, out["raw"][name].float(), f"step {step} {name}")
        H.assert_equal(below, out["masks"]["object_below_robot"], "object_below_robot")
        H.assert_equal(roll, out["masks"]["object_bad_orientation"], "bad_roll")
        # reset of done envs is the manager's job in per-term mode
        ids = out["done"].nonzero().flatten()
        if len(ids):
            oracle.gait.reset(ids)
            gait.reset(ids.to(cuda))
        H.assert_close(gait.valid_last_air_time, oracle.gait.vla, "gait state through the reference attribute name")
        synth.advance(env, keep_cmd_prob=1.0)
        denv2 = env.to(cuda)
        denv2.common_step_counter = denv.common_step_counter + 1
        setattr(denv2, "_locotouch_b200_fused", getattr(denv, "_locotouch_b200_fused"))
        denv = denv2


def test_object_state_term_and_binary_tactile_class(cuda, lt_lib):
    from locotouch_b200.mdp import observations as O
    from oracle.mdp import object_state_in_robot_frame as oracle_os

    env = synth.make_env(128, seed=9, with_object=True, with_tactile=True, tactile_jitter=0.2)
    denv = env.to(cuda)
    os_ = TS.ObjectStateObs()
    kw = dict(last_contact_time_threshold=os_.last_contact_time_threshold, current_contact_time_threshold=os_.current_contact_time_threshold,
              non_contact_obs=list(os_.non_contact_obs), n_min=list(os_.n_min), n_max=list(os_.n_max), scale=list(os_.scale))
    clean = O.object_state_in_robot_frame(denv, add_uniform_noise=False, **kw)
    H.assert_close(clean, oracle_os(env, os_, False), "object_state (clean)")
    noisy = O.object_state_in_robot_frame(denv, add_uniform_noise=True, **kw)
    assert noisy.shape == (128, 13) and not torch.equal(noisy, clean)
    assert float((noisy[:, :3] - clean[:, :3]).abs().max()) <= 0.0101
    # binary tactile class term
    params = dict(asset_cfg=SceneEntityCfg("robot", body_names="sensor_.*").resolve(denv.scene),
                  sensor_cfg=SceneEntityCfg("tactile_contact_sensor", body_names="sensor_.*").resolve(denv.scene), tactile_signal_shape=(17, 13),
                  contact_threshold=0.05, add_threshold_noise=True, threshold_n_min=-0.01, threshold_n_max=0.01, contact_dropout_prob=0.005,
                  contact_addition_prob=0.005, add_continuous_artifact=0.0)
    term = O.BinaryTactileSignals(SimpleNamespace(params=params), denv)
    thr = term.contact_threshold_envs_sensors
    assert float(thr.min()) >= 0.04 - 1e-6 and float(thr.max()) <= 0.06 + 1e-6  # 0.05 + U(-0.01, 0.01) in fp32
    g = torch.Generator().manual_seed(1)
    ud, ua = torch.rand(128, 221, generator=g), torch.rand(128, 221, generator=g)
    sig = term(denv, u_drop=ud.to(cuda), u_add=ua.to(cuda))
    ref = OT.binary_taxels(env.scene["robot"].data.body_quat_w[:, 17:], env.scene.sensors["tactile_contact_sensor"].data.net_forces_w,
                           thr.view(128, 221).cpu(), ud, ua)
    H.assert_equal(sig, ref["signal"], "BinaryTactileSignals.__call__")
    H.assert_equal(term.original_contact_taxels.view(128, 221), ref["original"], "original_contact_taxels")
    H.assert_equal(term.processed_contact_taxels.view(128, 221), ref["contact"], "processed_contact_taxels")
    assert term(denv).shape == (128, 442)  # production mode (Philox)


def test_tactile_recorder_and_action_term(cuda, lt_lib):
    from locotouch_b200.distill import TactileRecorder
    from locotouch_b200.mdp.actions import JointPositionActionPrevPrev

    n = 77
    rec = TactileRecorder(cuda, n, 442, min_delay=1, max_delay=3)
    ora = OT.TactileDelayOracle(n, 442, 1, 3, delay_steps=rec.delay_steps.cpu())
    assert int(rec.delay_steps.min()) >= 1 and int(rec.delay_steps.max()) <= 2  # exclusive high, like the reference
    g = torch.Generator().manual_seed(2)
    for step in range(6):
        x = (torch.rand(n, 442, generator=g) < 0.2).float()
        if step == 3:
            ids = torch.arange(0, n, 4)
            rec.reset(ids.to(cuda))
            ora.reset(ids, delay_steps=rec.delay_steps.cpu()[ids])
        rec.record_new_tactile_signals(x.to(cuda))
        ora.record(x)
        H.assert_equal(rec.get_tactile_signals(), ora.get(), f"delay line step {step}")
        H.assert_equal(rec.tactile_buffer, ora.buf, f"ring step {step}")
    # action term (reference actions.py:30-44)
    off = torch.randn(n, 12, generator=g)
    term = JointPositionActionPrevPrev(n, 12, cuda, scale=1.0, offset=off.to(cuda), clip_raw_actions=True, raw_action_clip_value=100.0, raw_action_scale=0.25)
    raw = torch.zeros(n, 12)
    prev = torch.zeros(n, 12)
    for step in range(3):
        a = torch.randn(n, 12, generator=g) * (300.0 if step == 1 else 1.0)
        term.process_actions(a.to(cuda))
        pprev, prev = prev, raw
        raw = torch.clamp(a, -100.0, 100.0) * 0.25
        H.assert_equal(term.raw_actions, raw, "raw_actions")
        H.assert_equal(term.prev_raw_actions, prev, "prev_raw_actions")
        H.assert_equal(term.prev_prev_raw_actions, pprev, "prev_prev_raw_actions")
        H.assert_equal(term.processed_actions, raw * 1.0 + off, "processed_actions")


def test_student_batch_helpers(cuda, lt_lib):
    from locotouch_b200.distill import masked_mse_loss, pad_trajectories

    g = torch.Generator().manual_seed(3)
    lengths = torch.tensor([5, 1, 9, 3, 9])
    offsets = torch.cat([torch.zeros(1, dtype=torch.long), lengths.cumsum(0)[:-1]])
    flat = torch.randn(int(lengths.sum()), 270, generator=g)
    out, masks = pad_trajectories(flat.to(cuda), offsets.to(cuda), lengths.to(cuda))
    ref = torch.zeros(9, 5, 270)
    ref_m = torch.zeros(9, 5, dtype=torch.bool)
    for b in range(5):  # reference replay_buffer.py:101-106
        ref[: lengths[b], b] = flat[offsets[b]: offsets[b] + lengths[b]]
        ref_m[: lengths[b], b] = True
    H.assert_equal(out, ref, "padded batch")
    H.assert_equal(masks, ref_m, "masks")
    s = torch.randn(9, 5, 12, generator=g, requires_grad=True)
    t = torch.randn(9, 5, 12, generator=g)
    loss_ref = (((s - t) ** 2).mean(-1) * ref_m).sum() / ref_m.sum()  # reference student.py:131,142
    loss_ref.backward()
    sd = s.detach().to(cuda).requires_grad_(True)
    loss = masked_mse_loss(sd, t.to(cuda), ref_m.to(cuda))
    loss.backward()
    H.assert_close(loss, loss_ref.detach(), "masked MSE")
    H.assert_close(sd.grad, s.grad, "masked MSE gradient", rtol=1e-5, atol=1e-8)


def test_action_term_matches_reference_golden(cuda, lt_lib):
    """The drop-in JointPositionActionPrevPrev (K0 kernel + reset) vs the reference class (tests/golden/action_term_c0.npz, generated
    from locotouch/mdp/actions.py:13-69): all six state tensors after every process_actions and every reset(env_ids), bit-exact."""
    from locotouch_b200.mdp.actions import JointPositionActionPrevPrev

    gold = H.load_golden("action_term_c0.npz")
    c = H.ACTION_TERM
    acts, offset, resets = H.action_term_tape()
    term = JointPositionActionPrevPrev(c["n"], c["J"], cuda, scale=c["scale"], offset=offset.to(cuda), clip_raw_actions=True,
                                       raw_action_clip_value=c["clip"], raw_action_scale=c["raw_scale"])
    names = ("raw_actions", "prev_raw_actions", "prev_prev_raw_actions", "processed_actions", "prev_processed_actions", "prev_prev_processed_actions")
    for s in range(c["steps"]):
        term.process_actions(acts[s].to(cuda))
        for k in names:
            H.assert_equal(getattr(term, k), gold[f"{k}_processed"][s], f"step {s} {k} after process_actions")
        term.reset(resets[s].to(cuda))
        for k in names:
            H.assert_equal(getattr(term, k), gold[f"{k}_reset"][s], f"step {s} {k} after reset")
