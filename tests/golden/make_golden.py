"""Generates the committed golden vectors by running the UNMODIFIED reference (``/root/reference``) on seeded inputs.

    PYTHONPATH=. python tests/golden/make_golden.py

Only runs in the build container (the reference is mounted read-only there; it does not exist on the GPU box).  The
inputs are NOT stored: they are regenerated from the same seeds by ``tests/helpers.py`` /
``locotouch_b200.sim.synth``; every fixture carries an input checksum so that generator drift is detected.

Fixtures (all < 1 MB):
  mdp_locomotion.npz / mdp_teacher.npz  per-step raw values of every reward term, custom terminations, object-state
                                        observation (clean + noisy) and the final gait state, from the reference
                                        ``locotouch/mdp`` callables driven through the IsaacLab stub namespace.
  ppo_c1.npz                            RolloutStorage.compute_returns outputs and one full PPO.update() (losses,
                                        learning-rate, parameters after the update) from the reference ``loco_rl``.
  commands_c3.npz                       velocity command term (multi-sampling bins, zero-command steps, gait logging metrics) and the
                                        reward-driven velocity curriculum over 70 steps of a reset tape, torch global RNG seeded.
  action_term_c0.npz                    JointPositionActionPrevPrev state after every process_actions / reset(env_ids) of a seeded tape.
  tactile_c4.npz                        BinaryTactileSignals bitmaps (explicit dropout / addition uniforms) and
                                        TactileRecorder outputs across resets.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.mdp import MdpOracle  # noqa: E402
from tests import helpers as H  # noqa: E402
from locotouch_b200.sim import synth  # noqa: E402
from locotouch_b200.sim.scene import SceneEntityCfg  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def env_checksum(env) -> float:
    return float(sum(t.double().abs().sum() for k, t in env.named_tensors().items() if k not in ("terminated", "time_outs")))


class patched_rand:
    """Feeds the reference's ``torch.rand_like`` / ``torch.rand`` calls from a queue of explicit tensors."""

    def __init__(self, rand_like_queue=(), rand_queue=()):
        self.like_q, self.rand_q = list(rand_like_queue), list(rand_queue)

    def __enter__(self):
        self._rl, self._r = torch.rand_like, torch.rand

        def rand_like(x, *a, **k):
            if self.like_q:
                item = self.like_q.pop(0)
                out = item(x) if callable(item) else item
                if out is not None:
                    assert out.shape == x.shape, (out.shape, x.shape)
                    return out.to(x.dtype)
            return self._rl(x, *a, **k)

        def rand(*a, **k):
            if self.rand_q:
                return self.rand_q.pop(0)
            return self._r(*a, **k)

        torch.rand_like, torch.rand = rand_like, rand
        return self

    def __exit__(self, *exc):
        torch.rand_like, torch.rand = self._rl, self._r


def reference_reward_terms(rw, env, spec, gait):
    """{name: raw tensor} from the reference callables, in cfg order, for the non-[IL] terms with non-zero weight."""
    sc = lambda *a, **k: SceneEntityCfg(*a, **k).resolve(env.scene)  # noqa: E731
    feet = sc("robot", body_names=".*foot")
    feet_s = sc("robot_contact_senosr", body_names=".*foot")
    out = {
        "track_lin_vel_xy": rw.track_lin_vel_xy_pst(env, 0.25, "base_velocity"),
        "track_ang_vel_z": rw.track_ang_vel_z_pst(env, 0.25, "base_velocity"),
        "foot_slip": rw.foot_slipping_ngt(env, 0.5, feet, feet_s),
        "foot_dragging": rw.foot_dragging_ngt(env, feet, 0.03, 0.1),
        "gait": gait(env, **gait.cfg.params),
        "track_base_height": rw.track_base_height_ngt(env, 0.42),
        "base_z_velocity": rw.base_z_velocity_ngt(env),
        "base_roll_pitch_angle": rw.base_roll_pitch_angle_ngt(env),
        "base_roll_pitch_velocity": rw.base_roll_pitch_velocity_ngt(env),
        "joint_position_limit": rw.joint_position_limit_ngt(env),
        "joint_position": rw.joint_position_ngt(env, SceneEntityCfg("robot"), 5.0, 0.3),
        "joint_acceleration": rw.joint_acceleration_ngt(env),
        "joint_velocity": rw.joint_velocity_ngt(env),
        "joint_torque": rw.joint_torque_ngt(env),
        "action_rate": rw.action_rate_ngt(env),
        "thigh_calf_collision": rw.thigh_calf_collision_ngt(env, 0.1, sc("robot_contact_senosr", body_names=[".*thigh", ".*calf"])),
    }
    if spec.with_object:
        out.update({
            "object_xy_position": rw.object_relative_xy_position_ngt(env, work_only_when_cmd=1),
            "object_xy_velocity": rw.object_relative_xy_velocity_ngt(env),
            "object_z_contact": rw.object_lose_contact_ngt(env, sensor_cfg=SceneEntityCfg("object_contact_sensor")),
            "object_z_velocity": rw.object_relative_z_velocity_ngt(env),
            "object_roll_pitch_angle": rw.object_relative_roll_angle_ngt(env),
            "object_roll_pitch_velocity": rw.object_relative_roll_velocity_ngt(env),
            "object_yaw_alignment": rw.object_relative_yaw_angle_ngt(env, work_only_when_cmd=1),
            "object_dangerous_state": rw.object_dangerous_state_ngt(env, x_max=0.125, y_max=0.097, z_min=0.095, roll_pitch_max=None, vel_xy_max=2.5),
            # extra: the two generic (non-cylinder) variants the object task can select (weights 0 in the cylinder cfg)
            "x_object_rp_angle": rw.object_relative_roll_pitch_angle_ngt(env),
            "x_object_rp_velocity": rw.object_relative_roll_pitch_velocity_ngt(env),
        })
    return {k: v.float() for k, v in out.items()}


def golden_mdp(scenario: str):
    rw, ob, te, _ = ref_loader.load_reference_mdp()
    spec, env, steps = H.make_mdp_env(scenario)
    oracle = MdpOracle(env, spec)  # supplies the [IL] termination set that decides which envs are reset
    cfg = H.gait_cfg(spec)
    gait_cls = rw.AdaptiveSymmetricGaitRewardwithObject if spec.with_object else rw.AdaptiveSymmetricGaitReward
    gait = gait_cls(cfg, env)
    names, raws, sums = None, [], []
    extra = {k: [] for k in ("bad_roll", "object_below_robot", "obj_state_clean", "obj_state_noisy", "done")}
    os_ = spec.object_state
    for step in range(steps):
        sums.append(env_checksum(env))
        masks, terminated, time_outs = oracle.terminations(env)
        vals = reference_reward_terms(rw, env, spec, gait)
        names = list(vals)
        raws.append(torch.stack([vals[k] for k in names]))
        done = terminated | time_outs
        extra["done"].append(done.clone())
        if spec.with_object:
            extra["bad_roll"].append(te.bad_roll(env, math.pi / 3, SceneEntityCfg("object")))
            extra["object_below_robot"].append(te.object_below_robot(env))
            kw = dict(last_contact_time_threshold=os_.last_contact_time_threshold, current_contact_time_threshold=os_.current_contact_time_threshold,
                      non_contact_obs=list(os_.non_contact_obs), n_min=list(os_.n_min), n_max=list(os_.n_max), scale=list(os_.scale))
            extra["obj_state_clean"].append(ob.object_state_in_robot_frame(env, add_uniform_noise=False, **kw))
            u_obs, u_euler = H.mdp_noise(scenario, step, env.num_envs, spec.obs_dim_per_step)
            u_state = u_obs[:, -13:]
            s = env.scene.sensors["object_contact_sensor"].data
            never = torch.logical_and(s.last_contact_time < os_.last_contact_time_threshold, s.current_contact_time < os_.current_contact_time_threshold).reshape(-1)
            with patched_rand([u_state.clone(), u_state[never].clone()], [u_euler.clone()]):
                extra["obj_state_noisy"].append(ob.object_state_in_robot_frame(env, add_uniform_noise=True, **kw))
        ids = done.nonzero(as_tuple=False).flatten()
        if len(ids) > 0:
            gait.reset(ids)
        H.advance_mdp_env(env, step)
    out = dict(
        term_names=np.array(names),
        raw=torch.stack(raws).numpy(),
        done=torch.stack(extra["done"]).numpy(),
        input_checksum=np.array(sums),
        gait_valid_last_air_time=gait.valid_last_air_time.numpy(),
        gait_swinging_in_zero_cmd=gait.swinging_in_zero_cmd.numpy(),
        gait_valid_previous_contact=gait.valid_previous_contact.numpy(),
        gait_last_velocity_cmd=gait.last_velocity_cmd.numpy(),
        gait_step_from_changing_cmd=gait.step_from_changing_cmd.numpy(),
        gait_last_step_current_air_time=gait.last_step_current_air_time.numpy(),
        gait_last_step_current_contact_time=gait.last_step_current_contact_time.numpy(),
    )
    if spec.with_object:
        for k in ("bad_roll", "object_below_robot", "obj_state_clean", "obj_state_noisy"):
            out[k] = torch.stack(extra[k]).numpy()
    np.savez_compressed(os.path.join(OUT, f"mdp_{scenario}.npz"), **out)
    frac_vla = float((gait.valid_last_air_time > 0.04).float().mean())
    print(f"mdp_{scenario}: raw {out['raw'].shape}, done rate {out['done'].mean():.3f}, gait mean {out['raw'][:, names.index('gait')].mean():.3f}, final VLA>0.04 {frac_vla:.2f}")


PPO_SMALL = dict(T=24, N=32, obs_dim=270, A=12, hidden=[64, 48, 32], seed=5)


def golden_ppo():
    ref_loader.load_reference_loco_rl()
    from loco_rl.algorithms import PPO
    from loco_rl.modules import ActorCritic
    from loco_rl.storage import RolloutStorage

    # ---- GAE alone on C1-shaped tensors (incl. the all-done and done-at-last-step envs)
    r = H.make_rollout(T=24, N=64, seed=0)
    st = RolloutStorage(64, 24, [270], [270], [12])
    st.rewards.copy_(r["rewards"])
    st.values.copy_(r["values"])
    st.dones.copy_(r["dones"].byte())
    st.compute_returns(r["last_values"], 0.99, 0.95, normalize_advantage=True)
    out = dict(gae_returns=st.returns.numpy().copy(), gae_advantages=st.advantages.numpy().copy())
    st.compute_returns(r["last_values"], 0.99, 0.95, normalize_advantage=False)
    out["gae_advantages_raw"] = st.advantages.numpy().copy()

    # ---- act -> process_env_step -> compute_returns -> update with the LocoTouch PPO cfg (rsl_rl_ppo_cfg.py:17-30)
    c = PPO_SMALL
    T, N, A = c["T"], c["N"], c["A"]
    torch.manual_seed(c["seed"])
    ac = ActorCritic(c["obs_dim"], c["obs_dim"], A, c["hidden"], c["hidden"], "elu", 1.0)
    init = torch.cat([p.detach().flatten() for p in ac.parameters()]).clone()
    alg = PPO(ac, num_learning_epochs=2, num_mini_batches=2, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
              entropy_coef=0.01, learning_rate=1.0e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
              desired_kl=0.01, device="cpu")
    alg.init_storage(N, T, [c["obs_dim"]], [c["obs_dim"]], [A])
    r = H.make_rollout(T=T, N=N, obs_dim=c["obs_dim"], A=A, seed=c["seed"])
    g = torch.Generator().manual_seed(c["seed"] + 1)
    eps = torch.randn(T, N, A, generator=g)
    from torch.distributions import Normal
    orig_sample = Normal.sample
    for t in range(T):
        Normal.sample = lambda self, sample_shape=torch.Size(), _e=eps[t]: (self.loc + self.scale * _e).detach()
        alg.act(r["obs"][t], r["critic_obs"][t])
        Normal.sample = orig_sample
        infos = {"time_outs": r["time_outs"][t, :, 0]}
        alg.process_env_step(r["rewards"][t, :, 0].clone(), r["dones"][t, :, 0].long(), infos)
    alg.compute_returns(r["critic_obs"][-1])
    out.update(
        ppo_init_params=init.numpy(),
        ppo_eps=eps.numpy(),
        ppo_actions=alg.storage.actions.numpy().copy(),
        ppo_logp=alg.storage.actions_log_prob.numpy().copy(),
        ppo_values=alg.storage.values.numpy().copy(),
        ppo_rewards=alg.storage.rewards.numpy().copy(),
        ppo_returns=alg.storage.returns.numpy().copy(),
        ppo_advantages=alg.storage.advantages.numpy().copy(),
    )
    perms = []
    orig_randperm = torch.randperm

    def randperm(n, *a, **k):
        p = orig_randperm(n, *a, **k)
        perms.append(p.clone())
        return p

    torch.randperm = randperm
    torch.manual_seed(c["seed"] + 2)
    # record the learning rate after every mini-batch
    lrs = []
    orig_step = alg.optimizer.step

    def step(*a, **k):
        lrs.append(alg.optimizer.param_groups[0]["lr"])
        return orig_step(*a, **k)

    alg.optimizer.step = step
    losses = alg.update()
    torch.randperm = orig_randperm
    out.update(
        ppo_perm=perms[0].numpy(),
        ppo_lr_sequence=np.array(lrs),
        ppo_losses=np.array([losses[0], losses[1], losses[2]]),
        ppo_final_params=torch.cat([p.detach().flatten() for p in ac.parameters()]).numpy(),
        ppo_param_names=np.array([k for k, _ in ac.named_parameters()]),
    )
    np.savez_compressed(os.path.join(OUT, "ppo_c1.npz"), **out)
    print(f"ppo_c1: losses {losses[:3]}, lr sequence {lrs}, |dparams| {float((torch.as_tensor(out['ppo_final_params']) - init).abs().mean()):.3e}")


TACTILE = dict(N=48, seed=21, steps=4)


def tactile_inputs(step: int, n: int, p_case: int):
    g = torch.Generator().manual_seed(1000 * (p_case + 1) + step)
    return torch.rand(n, 17, 13, generator=g), torch.rand(n, 17, 13, generator=g)


def golden_tactile():
    _, ob, _, _ = ref_loader.load_reference_mdp()
    _, rec_mod, _ = ref_loader.load_reference_distill()
    from types import SimpleNamespace

    c = TACTILE
    out = {}
    for p_case, (p_drop, p_add, jitter) in enumerate([(0.005, 0.005, 0.0), (0.2, 0.1, 0.3)]):
        env = synth.make_env(c["N"], seed=c["seed"] + p_case, with_object=True, with_tactile=True, tactile_jitter=jitter)
        params = dict(
            asset_cfg=SceneEntityCfg("robot", body_names="sensor_.*").resolve(env.scene),
            sensor_cfg=SceneEntityCfg("tactile_contact_sensor", body_names="sensor_.*").resolve(env.scene),
            tactile_signal_shape=(17, 13), contact_threshold=0.05, add_threshold_noise=True, threshold_n_min=-0.05 * 0.2,
            threshold_n_max=0.05 * 0.2, contact_dropout_prob=p_drop, contact_addition_prob=p_add, add_continuous_artifact=0.0,
            artifact_taxel_num_min=0, artifact_taxel_num_max=3, add_force_noise=True, force_n_prop_min=-0.1, force_n_prop_max=0.1,
            maximal_force=3.0, total_levels=5, add_level_noise=True, level_n_min=-1, level_n_max=1)
        g = torch.Generator().manual_seed(77 + p_case)
        u_thr = torch.rand(c["N"], 17, 13, generator=g)
        with patched_rand([u_thr.clone()]):
            term = ob.BinaryTactileSignals(SimpleNamespace(params=params), env)
        recorder = rec_mod.TactileRecorder("cpu", c["N"], 442, min_delay=1, max_delay=2)
        sigs, delayed = [], []
        for step in range(c["steps"]):
            u_drop, u_add = tactile_inputs(step, c["N"], p_case)
            with patched_rand([u_drop.clone(), None, u_add.clone()]):
                sig = term(env, **params)
            sigs.append(sig.clone())
            if step == 2:  # reset a third of the envs before recording, like ReplayBuffer.collect_data does on dones
                recorder.reset(torch.arange(0, c["N"], 3))
            recorder.record_new_tactile_signals(sig)
            delayed.append(recorder.get_tactile_signals().clone())
            synth.advance(env, tactile_jitter=jitter)
        out[f"thresholds_{p_case}"] = term.contact_threshold_envs_sensors.numpy().copy()
        out[f"signal_{p_case}"] = torch.stack(sigs).numpy().astype(np.uint8)
        out[f"delayed_{p_case}"] = torch.stack(delayed).numpy().astype(np.uint8)
        out[f"normal_forces_last_{p_case}"] = term.original_normal_forces.numpy().copy()
        print(f"tactile case {p_case}: contact fraction {out[f'signal_{p_case}'].mean():.4f}")
    np.savez_compressed(os.path.join(OUT, "tactile_c4.npz"), **out)


TACTILE_FORCE = dict(N=96, seed=23, steps=3)


def tactile_force_uniforms(step: int, n: int):
    g = torch.Generator().manual_seed(9000 + step)
    return {k: torch.rand(n, 221, generator=g) for k in ("drop", "drop_force", "add", "add_force", "noise", "small", "level")}


TACTILE_FORCE_PARAMS = dict(tactile_signal_shape=(17, 13), contact_threshold=0.05, add_threshold_noise=True, threshold_n_min=-0.01,
                            threshold_n_max=0.01, contact_dropout_prob=0.15, contact_addition_prob=0.05, add_continuous_artifact=0.0,
                            artifact_taxel_num_min=0, artifact_taxel_num_max=3, add_force_noise=True, force_n_prop_min=-0.6,
                            force_n_prop_max=0.1, maximal_force=1.5, total_levels=5, add_level_noise=True, level_n_min=-1, level_n_max=1)


def golden_tactile_forces():
    """Processed / Normalized / Discrete / Continuous tactile signal classes of the reference (observations.py:311-429) with
    every ``rand_like`` draw fed from explicit uniforms; the masked draws are cut out of full-size tensors with the masks the
    oracle derives (a wrong oracle mask would make the shapes disagree and the patched draw assert)."""
    from oracle import tactile as OT
    from types import SimpleNamespace

    _, ob, _, _ = ref_loader.load_reference_mdp()
    c = TACTILE_FORCE
    env = synth.make_env(c["N"], seed=c["seed"], with_object=True, with_tactile=True, tactile_jitter=0.2)
    params = dict(asset_cfg=SceneEntityCfg("robot", body_names="sensor_.*").resolve(env.scene),
                  sensor_cfg=SceneEntityCfg("tactile_contact_sensor", body_names="sensor_.*").resolve(env.scene), **TACTILE_FORCE_PARAMS)
    g = torch.Generator().manual_seed(555)
    u_thr = torch.rand(c["N"], 17, 13, generator=g)
    terms = {}
    for name in ("ProcessedTactileSignals", "NormalizedTactileSignals", "DiscreteTactileSignals", "CotinuousTactileSignals"):
        with patched_rand([u_thr.clone()]):
            terms[name] = getattr(ob, name)(SimpleNamespace(params=params), env)
    thr = terms["ProcessedTactileSignals"].contact_threshold_envs_sensors.reshape(c["N"], 221)
    out = dict(thresholds=thr.numpy().copy())
    sensor_ids = params["sensor_cfg"].body_ids
    for step in range(c["steps"]):
        u = tactile_force_uniforms(step, c["N"])
        quat = env.scene["robot"].data.body_quat_w[:, params["asset_cfg"].body_ids]
        force = env.scene.sensors["tactile_contact_sensor"].data.net_forces_w[:, sensor_ids]
        o = OT.force_signals(quat, force, thr, u, p_drop=0.15, p_add=0.05, add_force_noise=True, force_n_prop_min=-0.6, force_n_prop_max=0.1,
                             maximal_force=1.5, total_levels=5, add_level_noise=True, level_n_min=-1, level_n_max=1)
        sh = lambda t: t.reshape(c["N"], 17, 13)  # noqa: E731

        def queue():
            return [sh(u["drop"]).clone(), u["drop_force"][o["drop_mask"]].clone(), sh(u["add"]).clone(), u["add_force"][o["add_mask"]].clone(),
                    u["noise"][o["noise_mask"]].clone(), u["small"][o["small_mask"]].clone(), sh(u["level"]).clone()]

        with patched_rand(queue()):
            sig = terms["ProcessedTactileSignals"](env, **params)
        out[f"processed_{step}"] = sig.numpy().copy()
        if step == 0:
            print("drop / add / small counts:", int(o["drop_mask"].sum()), int(o["add_mask"].sum()), int(o["small_mask"].sum()),
                  " contact fraction", float(o["contact"].float().mean()))
            for name, key in (("NormalizedTactileSignals", "normalized_0"), ("DiscreteTactileSignals", "discrete_0"), ("CotinuousTactileSignals", "continuous_0")):
                q = queue()
                if name != "DiscreteTactileSignals":
                    q = q[:6]  # no discretisation -> no level-noise draw
                with patched_rand(q):
                    out[key] = terms[name](env, **params).numpy().copy()
        synth.advance(env, tactile_jitter=0.2)
    np.savez_compressed(os.path.join(OUT, "tactile_force_c4.npz"), **out)
    print("tactile_force_c4:", {k: v.shape for k, v in out.items() if k.endswith("_0")})


STUDENT_SMALL = dict(rnn_hidden=32, enc_hidden=[32, 16], pol_hidden=[32, 16], cnn_channels=(4, 4, 4), L=11, B=4, lengths=[11, 3, 7, 1], seed=31)


def student_batch(c=STUDENT_SMALL):
    g = torch.Generator().manual_seed(c["seed"])
    L, B = c["L"], c["B"]
    masks = torch.zeros(L, B, dtype=torch.bool)
    for b, n in enumerate(c["lengths"]):
        masks[:n, b] = True
    batch = dict(proprioceptions=torch.randn(L, B, 270, generator=g), teacher_encoder_obses=torch.randn(L, B, 78, generator=g),
                 tactile_signals=(torch.rand(L, B, 442, generator=g) < 0.15).float(), masks=masks)
    for k in ("proprioceptions", "teacher_encoder_obses", "tactile_signals"):
        batch[k] = batch[k] * masks.unsqueeze(-1)  # padded region is zero, as _prepare_padded_sequence leaves it
    teacher_w = torch.randn(12, 348, generator=g) * 0.05
    teacher_b = torch.randn(12, generator=g) * 0.05
    return batch, teacher_w, teacher_b


def shrink_student_cfg(cfg, c=STUDENT_SMALL):
    cfg.device = "cpu"
    cfg.tactile_encoder.rnn_hidden_size = c["rnn_hidden"]
    cfg.tactile_encoder.hidden_dims = list(c["enc_hidden"])
    cfg.student_policy.hidden_dims = list(c["pol_hidden"])
    cfg.pre_encoder.cnn_channels = tuple(c["cnn_channels"])
    return cfg


def golden_student():
    student_mod, _, cfg_mod = ref_loader.load_reference_distill()
    cfg = shrink_student_cfg(cfg_mod.DistillationRandCylinderCNNRNNMonCfg())
    batch, tw, tb = student_batch()
    teacher = lambda x: torch.nn.functional.linear(x, tw, tb)  # noqa: E731
    torch.manual_seed(STUDENT_SMALL["seed"])
    ref = student_mod.Student(cfg, 270, 442, 12, teacher_policy_inference=teacher)
    names = list(ref.state_dict().keys())
    init = torch.cat([v.flatten() for v in ref.state_dict().values()]).clone()
    # one step of the loop body of Student.train_on_data (reference student.py:121-151)
    ref._optimizer.zero_grad()
    actions = ref.forward(batch["proprioceptions"], batch["tactile_signals"])
    teacher_actions = teacher(torch.cat((batch["proprioceptions"], batch["teacher_encoder_obses"]), dim=-1))
    loss = ref._criterion(actions, teacher_actions).mean(dim=-1)
    loss = (loss * batch["masks"]).sum() / batch["masks"].sum()
    loss.backward()
    grad_norms = torch.stack([p.grad.norm() for p in ref.parameters()])
    ref._optimizer.step()
    after = torch.cat([v.flatten() for v in ref.state_dict().values()])
    np.savez_compressed(os.path.join(OUT, "student_c4.npz"), names=np.array(names), init=init.numpy(), actions=actions.detach().numpy(),
                        loss=np.array(float(loss)), grad_norms=grad_norms.numpy(), after=after.detach().numpy())
    print(f"student_c4: {init.numel()} params, loss {float(loss):.6f}")


STUDENT_CNN = dict(M=37, seed=53, E=64)


def student_cnn_inputs(c=STUDENT_CNN):
    """Frames for the full-geometry pre-encoder: binary taxel bitmaps (both channels equal, ~12 % set) and, second half, fp32
    frames with independent channels (the normalised / continuous encodings)."""
    g = torch.Generator().manual_seed(c["seed"])
    M = c["M"]
    bits = (torch.rand(M, 1, 17 * 13, generator=g) < 0.12).float().expand(M, 2, 17 * 13).reshape(M, 442).contiguous()
    dense = torch.rand(M, 442, generator=g)
    return bits, dense


def golden_student_cnn():
    """reference loco_rl/models/cnn_2d.py CNN2dHead at the LocoTouch student geometry (model_cfg.py:17-25), built by the reference's
    own generate_model from DistillationRandCylinderCNNRNNMonCfg.pre_encoder."""
    _, _, cfg_mod = ref_loader.load_reference_distill()
    loco = ref_loader.load_reference_loco_rl()
    from loco_rl.models.model_generation import generate_model  # the reference package (asserted by the loader)

    assert os.path.realpath(loco.__file__).startswith(os.path.realpath(ref_loader.REFERENCE_ROOT))
    cfg = cfg_mod.DistillationRandCylinderCNNRNNMonCfg()
    torch.manual_seed(STUDENT_CNN["seed"])
    net = generate_model(442, cfg.pre_encoder.embedding_dim, cfg.pre_encoder)
    names = list(net.state_dict().keys())
    flat = torch.cat([v.flatten() for v in net.state_dict().values()]).clone()
    bits, dense = student_cnn_inputs()
    with torch.no_grad():
        out_bits = net(bits.reshape(-1, 2, 17, 13))
        out_dense = net(dense.reshape(-1, 2, 17, 13))
    np.savez_compressed(os.path.join(OUT, "student_cnn_c4.npz"), names=np.array(names), params=flat.numpy(), out_bits=out_bits.numpy(),
                        out_dense=out_dense.numpy(), checksum=np.array(float(bits.double().sum() + dense.double().sum())))
    print(f"student_cnn_c4: {flat.numel()} params, out {tuple(out_bits.shape)}")


RECURRENT_SMALL = dict(T=24, N=48, D=10, A=3, L=2, Hd=6, seed=11, num_mini_batches=4)


def recurrent_inputs(c=RECURRENT_SMALL):
    g = torch.Generator().manual_seed(c["seed"])
    T, N = c["T"], c["N"]
    dones = (torch.rand(T, N, 1, generator=g) < 0.08)
    dones[:, 0] = True   # an env that is done at every step
    dones[:, 1] = False  # one that never is (the forced cut at the last step makes a single trajectory)
    dones[-1, 2] = True  # done exactly at the last step
    dones[0, 3] = True   # done at the first step
    return dict(obs=torch.randn(T, N, c["D"], generator=g), critic_obs=torch.randn(T, N, c["D"] + 2, generator=g), dones=dones,
                hid_a=torch.randn(T, c["L"], N, c["Hd"], generator=g), hid_c=torch.randn(T, c["L"], N, c["Hd"], generator=g),
                actions=torch.randn(T, N, c["A"], generator=g))


def golden_recurrent():
    """utils.split_and_pad_trajectories / unpad_trajectories and RolloutStorage.recurrent_mini_batch_generator of the reference."""
    ref_loader.load_reference_loco_rl()
    from loco_rl.storage import RolloutStorage
    from loco_rl.utils import split_and_pad_trajectories, unpad_trajectories

    c = RECURRENT_SMALL
    r = recurrent_inputs(c)
    padded, masks = split_and_pad_trajectories(r["obs"], r["dones"])
    back = unpad_trajectories(padded, masks)
    assert torch.equal(back, r["obs"])
    st = RolloutStorage(c["N"], c["T"], [c["D"]], [c["D"] + 2], [c["A"]])
    for t in range(c["T"]):
        tr = RolloutStorage.Transition()
        tr.observations, tr.critic_observations, tr.actions = r["obs"][t], r["critic_obs"][t], r["actions"][t]
        tr.rewards, tr.dones, tr.values = torch.zeros(c["N"]), r["dones"][t, :, 0], torch.zeros(c["N"], 1)
        tr.actions_log_prob, tr.action_mean, tr.action_sigma = torch.zeros(c["N"]), r["actions"][t], r["actions"][t].abs()
        tr.hidden_states = (r["hid_a"][t], r["hid_c"][t])
        st.add_transitions(tr)
    out = dict(padded=padded.numpy(), masks=masks.numpy())
    for i, b in enumerate(st.recurrent_mini_batch_generator(c["num_mini_batches"], num_epochs=1)):
        obs_b, cobs_b, act_b, _, _, _, _, _, _, (hid_a, hid_c), masks_b, _ = b
        out[f"mb{i}_obs"], out[f"mb{i}_cobs"], out[f"mb{i}_actions"] = obs_b.numpy(), cobs_b.numpy(), act_b.numpy()
        out[f"mb{i}_hid_a"], out[f"mb{i}_hid_c"], out[f"mb{i}_masks"] = hid_a.numpy(), hid_c.numpy(), masks_b.numpy()
    np.savez_compressed(os.path.join(OUT, "recurrent_c1.npz"), **out)
    print(f"recurrent_c1: {padded.shape[1]} trajectories from {c['N']} envs x {c['T']} steps")


DAGGER_SMALL = dict(N=24, steps=200, P=22, tactile=16, first=150, second=120, batch=7, np_seed=3)


class _DummyStudent:
    """Stands in for the student during data collection: ignores its inputs, counts reset calls."""

    def __init__(self, num_envs, device="cpu"):
        self.num_envs, self.device, self.resets = num_envs, device, 0

    def __call__(self, proprioception, tactile):
        return torch.zeros(self.num_envs, 12, device=self.device)

    def reset(self, dones):
        self.resets += 1


RECURRENT_PPO = dict(T=24, N=32, A=12, flat_dim=30, enc_dim=18, hidden=[32, 24], enc_hidden=[24], pre_hidden=[20, 16], pre_emb=12, emb=10,
                     rnn_hidden=16, seed=77, num_learning_epochs=2, num_mini_batches=2)


def recurrent_policy_kwargs(kind: str, c=RECURRENT_PPO):
    """Constructor arguments of the two recurrent encoder policies (cfg blocks rsl_rl_ppo_cfg.py:147,181-226 at reduced widths)."""
    obs_dim = c["flat_dim"] + c["enc_dim"]
    kw = dict(actor_obs_dim=obs_dim, critic_obs_dim=obs_dim, num_actions=c["A"], actor_flatten_obs_end_idx=-c["enc_dim"],
              actor_encoder_obs_start_idx=-c["enc_dim"], actor_encoder_hidden_dims=list(c["enc_hidden"]), actor_encoder_embedding_dim=c["emb"],
              actor_hidden_dims=list(c["hidden"]), critic_flatten_obs_end_idx=-c["enc_dim"], critic_encoder_obs_start_idx=-c["enc_dim"],
              critic_encoder_hidden_dims=list(c["enc_hidden"]), critic_encoder_embedding_dim=c["emb"], critic_hidden_dims=list(c["hidden"]),
              encoder_rnn_type="gru", encoder_rnn_hidden_size=c["rnn_hidden"], encoder_rnn_num_layers=1, activation="elu", init_noise_std=1.0)
    if kind == "pre_rnn":
        kw.update(actor_pre_encoder_hidden_dims=list(c["pre_hidden"]), actor_pre_encoder_embedding_dim=c["pre_emb"],
                  critic_pre_encoder_hidden_dims=list(c["pre_hidden"]), critic_pre_encoder_embedding_dim=c["pre_emb"])
    return kw


def golden_recurrent_ppo():
    """One full PPO iteration (act x T with the hidden-state bookkeeping -> process_env_step -> compute_returns -> recurrent update) of
    the UNMODIFIED reference with ActorCriticRNNEncoder and ActorCriticPreEncoderRNNEncoder
    (loco_rl/modules/actor_critic_rnn_encoder.py, actor_critic_pre_encoder_rnn_encoder.py; algorithms/ppo.py:130,195-196,251-255)."""
    import contextlib
    import io

    ref_loader.load_reference_loco_rl()
    from loco_rl.algorithms import PPO
    from loco_rl.modules import ActorCriticPreEncoderRNNEncoder, ActorCriticRNNEncoder
    from torch.distributions import Normal

    c = RECURRENT_PPO
    T, N, A = c["T"], c["N"], c["A"]
    obs_dim = c["flat_dim"] + c["enc_dim"]
    out = {}
    for kind, cls in (("rnn", ActorCriticRNNEncoder), ("pre_rnn", ActorCriticPreEncoderRNNEncoder)):
        torch.manual_seed(c["seed"])
        with contextlib.redirect_stdout(io.StringIO()):
            ac = cls(**recurrent_policy_kwargs(kind))
        names = [k for k, _ in ac.named_parameters()]
        init = torch.cat([p.detach().flatten() for p in ac.parameters()]).clone()
        alg = PPO(ac, num_learning_epochs=c["num_learning_epochs"], num_mini_batches=c["num_mini_batches"], clip_param=0.2, gamma=0.99, lam=0.95,
                  value_loss_coef=1.0, entropy_coef=0.01, learning_rate=1.0e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
                  desired_kl=0.01, device="cpu")
        alg.init_storage(N, T, [obs_dim], [obs_dim], [A])
        r = H.make_rollout(T=T, N=N, obs_dim=obs_dim, A=A, seed=c["seed"])
        eps = torch.randn(T, N, A, generator=torch.Generator().manual_seed(c["seed"] + 1))
        orig_sample = Normal.sample
        with torch.inference_mode():  # the rollout of OnPolicyRunner.learn (runners/on_policy_runner.py) runs under inference_mode
            for t in range(T):
                Normal.sample = lambda self, sample_shape=torch.Size(), _e=eps[t]: (self.loc + self.scale * _e).detach()
                alg.act(r["obs"][t], r["critic_obs"][t])
                Normal.sample = orig_sample
                alg.process_env_step(r["rewards"][t, :, 0].clone(), r["dones"][t, :, 0].long(), {"time_outs": r["time_outs"][t, :, 0]})
            alg.compute_returns(r["critic_obs"][-1])
        st = alg.storage
        out.update({f"{kind}_names": np.array(names), f"{kind}_init": init.numpy(), f"{kind}_eps": eps.numpy(), f"{kind}_actions": st.actions.numpy().copy(),
                    f"{kind}_logp": st.actions_log_prob.numpy().copy(), f"{kind}_values": st.values.numpy().copy(),
                    f"{kind}_returns": st.returns.numpy().copy(), f"{kind}_advantages": st.advantages.numpy().copy(),
                    f"{kind}_hid_a": st.saved_hidden_states_a[0].numpy().copy(), f"{kind}_hid_c": st.saved_hidden_states_c[0].numpy().copy(),
                    f"{kind}_final_hidden": ac.get_hidden_states()[0].detach().numpy().copy()})
        lrs = []
        orig_step = alg.optimizer.step

        def step(*a, _orig=orig_step, _alg=alg, **k):
            lrs.append(_alg.optimizer.param_groups[0]["lr"])
            return _orig(*a, **k)

        alg.optimizer.step = step
        losses = alg.update()
        out.update({f"{kind}_lr_sequence": np.array(lrs), f"{kind}_losses": np.array([losses[0], losses[1], losses[2]]),
                    f"{kind}_final": torch.cat([p.detach().flatten() for p in ac.parameters()]).numpy()})
        print(f"recurrent_ppo[{kind}]: {init.numel()} params, losses {losses[:3]}, lr {lrs}")
    np.savez_compressed(os.path.join(OUT, "recurrent_ppo_c1.npz"), **out)


def golden_dagger():
    """ReplayBuffer.collect_data (teacher roll-out, then a student roll-out appended to it), to_recurrent_generator and
    evaluate of the reference, driven by the deterministic tape env of tests/helpers.py."""
    _, recorder_mod, _ = ref_loader.load_reference_distill()
    import sys as _sys
    rb_mod = _sys.modules["locotouch.distill.replay_buffer"]
    rb_mod.tqdm = lambda *a, **k: type("P", (), {"update": lambda self, n: None})()  # silence the progress bar
    c = DAGGER_SMALL
    env = H.TapeEnv(c["N"], c["steps"], obs_dim=c["P"] + 8, tactile_dim=c["tactile"])
    rec = recorder_mod.TactileRecorder("cpu", c["N"], c["tactile"], 1, 2)
    rb = rb_mod.ReplayBuffer(env, rec, c["P"])
    teacher = lambda x: torch.zeros(c["N"], 12)  # noqa: E731
    r1, l1 = rb.collect_data(teacher, None, c["first"])
    out = dict(rewards1=np.array(r1), lengths1=np.array(l1), steps1=np.array(rb.num_steps), trajs1=np.array(rb.num_trajs), t1=np.array(env.t))
    r2, l2 = rb.collect_data(teacher, _DummyStudent(c["N"]), c["second"])
    out.update(rewards2=np.array(r2), lengths2=np.array(l2), steps2=np.array(rb.num_steps), trajs2=np.array(rb.num_trajs), t2=np.array(env.t))
    out["traj_lengths"] = np.array([p.shape[0] for p in rb._proprioceptions])
    out["flat_prop"] = torch.cat(rb._proprioceptions).numpy()
    out["flat_teacher"] = torch.cat(rb._teacher_encoder_obses).numpy()
    out["flat_tactile"] = torch.cat(rb._tactile_signals).numpy()
    np.random.seed(c["np_seed"])
    batches = list(rb.to_recurrent_generator(c["batch"]))
    out["num_batches"] = np.array(len(batches))
    for i in (0, len(batches) - 1):
        b = batches[i]
        out[f"b{i}_prop"], out[f"b{i}_teacher"] = b["proprioceptions"].numpy(), b["teacher_encoder_obses"].numpy()
        out[f"b{i}_tactile"], out[f"b{i}_masks"] = b["tactile_signals"].numpy(), b["masks"].numpy()
    # evaluation on a fresh RSL-style tape
    env2 = H.TapeEnv(c["N"], c["steps"], seed=5, obs_dim=c["P"] + 8, tactile_dim=c["tactile"], rsl_style=True)
    rb2 = rb_mod.ReplayBuffer(env2, rec, c["P"])
    er, el = rb2.evaluate(_DummyStudent(c["N"]), 40)
    out["eval_rewards"], out["eval_lengths"] = np.array(er), np.array(el)
    np.savez_compressed(os.path.join(OUT, "dagger_c4.npz"), **out)
    print(f"dagger_c4: {rb.num_trajs} trajectories / {rb.num_steps} steps, {len(batches)} batches; eval {len(er)} episodes")


def reference_command_view(cmd, cur):
    def view():
        r, p = cmd.cfg.ranges, cmd.cfg.previous_ranges
        return dict(cmd=cmd.vel_command_b, buffer=cmd.vel_command_b_buffer, time_left=cmd.time_left, standing=cmd.is_standing_env,
                    counter=cmd.command_counter, metrics=cmd.metrics, ranges=[*r.lin_vel_x, *r.lin_vel_y, *r.ang_vel_z],
                    previous=[*p.lin_vel_x, *p.lin_vel_y, *p.ang_vel_z],
                    equal=[cmd.lin_vel_x_equal_ranges, cmd.lin_vel_y_equal_ranges, cmd.ang_vel_z_equal_ranges], izcs=cmd.initial_zero_command_steps,
                    rel_standing=cmd.cfg.rel_standing_envs,
                    curriculum=[cur.lin_forward_bins, cur.ang_forward_bins, cur.success_repeat_times_lin, cur.success_repeat_times_ang,
                                int(cur.env_reseted_lin.sum()), int(cur.env_reseted_ang.sum()), float(cur.episode_length_buf_lin.sum()),
                                float(cur.episode_length_buf_ang.sum())])
    return view


def golden_commands():
    """UniformVelocityCommandGaitLoggingMultiSampling + ModifyVelCommandsRangeBasedonReward (reference commands.py:427-576,
    curriculums.py:184-274, UNMODIFIED, over the restated [IL] base classes of oracle/il_commands.py) driven by
    tests/scenarios.CommandScenario under torch.manual_seed; the oracle reproduces the run call for call (TorchRng)."""
    from types import SimpleNamespace

    from tests import scenarios as S

    commands, curriculums = ref_loader.load_reference_commands()
    out = {}
    for tag, binary in (("", False), ("bin_", True)):
        torch.manual_seed(S.CMD_SCENARIO["seed"])
        sc = S.CommandScenario(binary_maximal_command=binary)
        c = S.CMD_CFG
        Cfg = commands.UniformVelocityCommandGaitLoggingMultiSamplingCfg
        cfg = Cfg(asset_name="robot", resampling_time_range=c["resampling_time_range"], rel_heading_envs=0.0, heading_command=False,
                  ranges=Cfg.Ranges(**c["ranges"]), new_command_probs=c["new_command_probs"], rel_standing_envs=c["rel_standing_envs"],
                  final_rel_standing_envs=c["final_rel_standing_envs"], initial_zero_command_steps=c["initial_zero_command_steps"],
                  final_initial_zero_command_steps=c["final_initial_zero_command_steps"], binary_maximal_command=binary,
                  sensor_cfg=SceneEntityCfg("robot_contact_senosr", body_names=".*foot"))
        cmd = commands.UniformVelocityCommandGaitLoggingMultiSampling(cfg, sc.env)
        assert list(cmd.sensor_cfg.body_ids) == S.FEET_SENSOR_IDS
        sc.env.command_manager._terms["base_velocity"] = cmd
        params = dict(command_name="base_velocity", reward_name_lin="track_lin_vel_xy", reward_name_ang="track_ang_vel_z", **S.CUR_CFG)
        cur = curriculums.ModifyVelCommandsRangeBasedonReward(SimpleNamespace(params=params), sc.env)
        rec = S.CommandRecorder(reference_command_view(cmd, cur))
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):  # set_ranges prints banners
            # binary_maximal_command: the curriculum stays out -- once it has run, cfg.ranges hold np.float64 values (np.clip,
            # curriculums.py:239) and commands.py:520-521 builds a float64 tensor that index_put_ rejects (latent reference bug)
            sc.run(cmd, (lambda env, ids: None) if binary else (lambda env, ids: cur(env, ids, **params)), rec)
        st = rec.stacked()
        for k, v in st.items():
            out[tag + k] = v.numpy()
        if not binary:
            assert st["scalars"][-1, 12:15].tolist() == [1.0, 1.0, 1.0] and st["scalars"][-1, 15] == c["final_initial_zero_command_steps"], "scenario must reach the final ranges"
    out["checksum"] = np.float64(env_checksum(S.CommandScenario().env))
    np.savez_compressed(os.path.join(OUT, "commands_c3.npz"), **out)


# ------------------------------------------------------------------------------------------------ action term (mdp/actions.py)
ACTION_STATE = ("raw_actions", "prev_raw_actions", "prev_prev_raw_actions", "processed_actions", "prev_processed_actions",
                "prev_prev_processed_actions")


def golden_action_term():
    """The reference ``JointPositionActionPrevPrev`` (locotouch/mdp/actions.py:13-69, unmodified) over a restated [IL] base: IsaacLab's
    JointAction keeps ``_raw_actions`` / ``_processed_actions``, ``process_actions`` stores the actions and applies scale + offset,
    ``reset`` zeroes the raw actions of the given envs.  State after every process_actions call and after every reset(env_ids)."""
    _, _, _, actions_mod = ref_loader.load_reference_mdp()
    base = sys.modules["isaaclab.envs.mdp.actions"].JointPositionAction
    c = H.ACTION_TERM

    def il_init(self, cfg, env):  # [IL] JointAction.__init__ / JointPositionAction.__init__ (use_default_offset: default joint positions)
        self.cfg, self._env = cfg, env
        self._raw_actions = torch.zeros(env.num_envs, cfg.action_dim)
        self._processed_actions = torch.zeros_like(self._raw_actions)
        self._scale, self._offset = cfg.scale, cfg.offset

    def il_process(self, actions):  # [IL] JointAction.process_actions
        self._raw_actions[:] = actions
        self._processed_actions = self._raw_actions * self._scale + self._offset

    def il_reset(self, env_ids=None):  # [IL] JointAction.reset
        self._raw_actions[env_ids] = 0.0

    base.__init__, base.process_actions, base.reset = il_init, il_process, il_reset
    base.raw_actions = property(lambda self: self._raw_actions)
    base.processed_actions = property(lambda self: self._processed_actions)
    acts, offset, resets = H.action_term_tape()
    from types import SimpleNamespace

    cfg = SimpleNamespace(clip_raw_actions=True, raw_action_clip_value=c["clip"], raw_action_scale=c["raw_scale"], scale=c["scale"],
                          offset=offset, action_dim=c["J"])
    term = actions_mod.JointPositionActionPrevPrev(cfg, SimpleNamespace(num_envs=c["n"]))
    out = {f"{k}_{when}": [] for k in ACTION_STATE for when in ("processed", "reset")}
    for s in range(c["steps"]):
        term.process_actions(acts[s])
        for k in ACTION_STATE:
            out[f"{k}_processed"].append(getattr(term, k).clone().numpy())
        term.reset(resets[s])
        for k in ACTION_STATE:
            out[f"{k}_reset"].append(getattr(term, k).clone().numpy())
    np.savez_compressed(os.path.join(OUT, "action_term_c0.npz"), input_checksum=np.float64(float(acts.double().sum() + offset.double().sum())),
                        num_reset=np.array([len(r) for r in resets]), **{k: np.stack(v) for k, v in out.items()})


if __name__ == "__main__":
    assert ref_loader.reference_available(), "the reference is not mounted"
    torch.set_num_threads(1)
    golden_mdp("locomotion")
    golden_mdp("teacher")
    golden_ppo()
    golden_tactile()
    golden_student()
    golden_recurrent()
    golden_dagger()
    golden_tactile_forces()
    golden_commands()
    golden_student_cnn()
    golden_recurrent_ppo()
    golden_action_term()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")

