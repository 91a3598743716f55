"""ORACLE (test infrastructure, never imported by the product path).

CPU (torch fp32) restatement of the per-step MDP manager terms of LocoTouch and of the IsaacLab manager semantics
around them.  Each function cites the reference lines it follows.  Pinned against the live reference by
tests/test_oracle_vs_reference.py (build container) and against tests/golden/mdp_*.npz (everywhere).

[IL] parts (RewardManager accumulation, TerminationManager OR, ObservationManager noise->scale->history, built-in
terms is_alive / time_out / bad_orientation / root_height_below_minimum / illegal_contact) restate IsaacLab 2.2.1
behaviour from SURVEY.md App. B; IsaacLab is not under /root/reference, so for those parts parity is UNPINNED.
"""
from __future__ import annotations

import math

import torch

from . import il_math as M

# kind codes duplicated on purpose (the oracle must not depend on product code for its arithmetic)
(ALIVE, LIN, ANG, SLIP, DRAG, GAIT, HEIGHT, ZVEL, RPANG, RPVEL, JLIM, JPOS, JACC, JVEL, JTORQ, ARATE, COLL,
 OXY, OXYVEL, OLOSE, OZVEL, ORPANG, ORPVEL, OROLL, OROLLVEL, OYAW, ODANGER) = range(27)
T_TIMEOUT, T_ORIENT, T_HEIGHT, T_CONTACT, T_BELOW, T_ROLL = range(6)


def _cmd(env):
    return env.command_manager.get_command("base_velocity")


def _contact_mask(env, body_ids, threshold):
    """max over the force history of |F| > threshold  (reference rewards.py:39-40, 464-465; [IL] illegal_contact)."""
    f = env.scene.sensors["robot_contact_senosr"].data.net_forces_w_history
    return torch.max(torch.linalg.norm(f[:, :, body_ids], dim=-1), dim=1)[0] > threshold


# ------------------------------------------------------------------------------------------------ gait (rewards.py:60-392)
class GaitOracle:
    """AdaptiveSymmetricGaitReward / ...withObject: SURVEY.md App. A.2 restated; state arrays as rewards.py:96-105."""

    def __init__(self, env, gp, feet_ids):
        n = env.num_envs
        self.gp = gp
        self.feet = list(feet_ids)  # (pair0[0], pair0[1], pair1[0], pair1[1]) in sensor body ids
        self.lsa = torch.zeros(n, 4)  # last_step_current_air_time
        self.lsc = torch.zeros(n, 4)  # last_step_current_contact_time
        self.sz = torch.zeros(n, 4, dtype=torch.bool)  # swinging_in_zero_cmd
        self.vla = torch.zeros(n, 4)  # valid_last_air_time
        self.vpc = torch.zeros(n, 4, dtype=torch.bool)  # valid_previous_contact
        self.last_cmd = torch.zeros(n, 3)
        self.steps = torch.zeros(n)
        self.theta = gp.judge_time_threshold
        self.slope = gp.rwd_upper_bound / (1.0 / (gp.soft_minimum_frequency * 2.0))  # rewards.py:72,78

    def reset(self, ids):
        """rewards.py:107-114"""
        for t in (self.lsa, self.lsc, self.vla, self.last_cmd, self.steps):
            t[ids] = 0.0
        self.sz[ids] = False
        self.vpc[ids] = False

    def update_state(self, env):
        """rewards.py:158-200, same statement order."""
        d = env.scene.sensors["robot_contact_senosr"].data
        a, c, la = d.current_air_time[:, self.feet], d.current_contact_time[:, self.feet], d.last_air_time[:, self.feet]
        cmd = _cmd(env)
        nz = torch.norm(cmd, dim=1) > 0.0
        th = self.theta
        self.vla[~nz, :] = 0.0
        new_swing = (self.lsa < th) & (a > th)
        self.sz[new_swing & nz.unsqueeze(-1)] = False
        self.sz[(a > th) & (~nz).unsqueeze(-1)] = True
        self.steps += 1
        chg = torch.any(torch.abs(cmd - self.last_cmd) > 1.0e-3, dim=1)
        self.last_cmd[chg] = cmd[chg].clone()
        self.steps[chg] = 0
        self.sz[chg] = True
        self.vla[chg, :] = 0.0
        if torch.any(nz):  # cross-env coupling, rewards.py:190
            new_land = (self.lsc < th) & (c > th)
            ok = new_land & self.vpc & (~self.sz)
            self.vla[ok] = la[ok].clone()
        self.lsa[:] = a
        self.lsc[:] = c
        self.vpc[c > th] = True

    def vel_score(self, env):
        """rewards.py:202-213"""
        cmd = _cmd(env)
        r = env.scene["robot"].data
        nz = torch.norm(cmd, dim=1) > 0.0
        e_lin = torch.linalg.norm(cmd[:, :2] - r.root_lin_vel_b[:, :2], dim=1)
        e_ang = torch.abs(cmd[:, 2] - r.root_ang_vel_b[:, 2])
        e_lin = torch.where(nz, e_lin, 0.0)
        e_ang = torch.where(nz, e_ang, 0.0)
        s = self.gp.vel_tracking_exp_sigma
        return (torch.exp(-e_lin / s) + torch.exp(-e_ang / s)) / 2.0

    def task_score(self, env):
        """rewards.py:215-216 (plain) / 372-392 (with object)."""
        s = self.vel_score(env)
        if not self.gp.with_object:
            return s
        r, o = env.scene["robot"].data, env.scene["object"].data
        _, _, yaw = M.euler_xyz_from_quat(r.root_quat_w)
        zero = torch.zeros_like(yaw)
        q_yaw = M.quat_from_euler_xyz(zero, zero, yaw)
        rel = M.quat_apply_inverse(q_yaw, o.root_pos_w - r.root_pos_w)
        xy = torch.abs(rel[:, :2])
        bx = torch.clip(1.0 - xy[:, 0] / self.gp.obj_x_max, min=0.0, max=1.0)
        by = torch.clip(1.0 - xy[:, 1] / self.gp.obj_y_max, min=0.0, max=1.0)
        bal = (bx + by) / 2.0
        return torch.clip((s * 2 + bal) / 3.0, min=0.0, max=1.0)

    def swing_bonus(self, env, k0, k1):
        """rewards.py:243-346; k0,k1 index the 4-foot gait order.  Select semantics over inf/NaN lanes preserved."""
        gp, th = self.gp, self.theta
        d = env.scene.sensors["robot_contact_senosr"].data
        a = d.current_air_time[:, [self.feet[k0], self.feet[k1]]]
        m = torch.mean(a, dim=1)
        both_air = torch.all(a > th, dim=1)
        tgt = [0, 1] if k0 in (0, 1) else [2, 3]
        oth = [2, 3] if tgt == [0, 1] else [0, 1]
        vt, vo = self.vla[:, tgt], self.vla[:, oth]
        m_t, m_o = torch.mean(vt, dim=1), torch.mean(vo, dim=1)
        two_dt = 2 * env.step_dt
        ok_t = torch.all(vt > th, dim=1) & torch.all(vt > two_dt, dim=1)
        ok_o = torch.all(vo > th, dim=1) & torch.all(vo > two_dt, dim=1)
        e = both_air & (ok_t | ok_o)
        ref = torch.where(e, m_o, 0.0)
        tol = ref + gp.tolerance_proportion * ref
        diff = torch.where(e, m_t - m_o, 0.0)
        ext = torch.clamp(tol - diff, min=ref, max=tol)
        within = e & (m <= ext)
        between = e & (m > ext) & (m <= tol)
        within = within | (e & (diff < 0.0))
        ub, lb = gp.rwd_upper_bound, gp.rwd_lower_bound
        r_within = torch.clamp(self.slope * m, max=ub)
        r_ref = torch.clamp(self.slope * ref, max=ub)
        r_ext = torch.clamp(self.slope * ext, max=ub)
        r_tol = torch.clamp(self.slope * tol, max=ub)
        lt = e & (ext < tol)
        a_b = torch.where(lt, -r_ext / (tol - ext), 0.0)
        b_b = torch.where(lt, -a_b * tol, 0.0)
        r_between = torch.where(lt, a_b * m + b_b, r_ext)
        gt = e & (ext > ref)
        a_y = torch.where(gt, -r_ref / (ext - ref), 0.0)
        b_y = torch.where(gt, -a_y * tol, 0.0)
        low = torch.where(lt, (diff / (gp.tolerance_proportion * ref)) * lb, r_tol)
        low[e & (~ok_o)] = lb
        low = torch.clamp(low, min=lb, max=ub)
        r_beyond = torch.where(gt, a_y * m + b_y, low)
        r_beyond = torch.clamp(r_beyond, min=low)
        r = torch.where(within, r_within, torch.where(between, r_between, r_beyond))
        return torch.where(e, r, 0.0)

    def sync(self, env, k0, k1, score):
        """rewards.py:218-241"""
        gp, th = self.gp, self.theta
        d = env.scene.sensors["robot_contact_senosr"].data
        f0, f1 = self.feet[k0], self.feet[k1]
        a = d.current_air_time[:, [f0, f1]]
        both_air = torch.all((a > th) & (a < gp.air_time_gait_bound), dim=1)
        c = d.current_contact_time
        c0 = (c[:, f0] > th) & (c[:, f0] < gp.contact_time_gait_bound)
        c1 = (c[:, f1] > th) & (c[:, f1] < gp.contact_time_gait_bound)
        both_contact = c0 & c1
        if gp.encourage_symmetricity_and_low_frequency > 0.5:
            bonus = self.swing_bonus(env, k0, k1)
            scale = 1 - gp.task_performance_ratio + gp.task_performance_ratio * score
            pos = bonus > 0.0
            bonus[pos] *= scale[pos]
            bonus += 1.0
            return torch.where(both_air, bonus, torch.where(both_contact, 1.0, 0.0))
        return torch.where(both_air | both_contact, 1.0, 0.0)

    def async_(self, env, k0, k1):
        """rewards.py:348-363"""
        gp, th = self.gp, self.theta
        d = env.scene.sensors["robot_contact_senosr"].data
        f0, f1 = self.feet[k0], self.feet[k1]
        a, c = d.current_air_time, d.current_contact_time
        th_async = th + gp.async_time_tolerance
        both = (c[:, f0] > th) & (c[:, f0] <= th_async) & (c[:, f1] > th) & (c[:, f1] <= th_async)
        a0 = (a[:, f0] > th) & (a[:, f0] < gp.air_time_gait_bound)
        a1 = (a[:, f1] > th) & (a[:, f1] < gp.air_time_gait_bound)
        c0 = (c[:, f0] > th) & (c[:, f0] < gp.contact_time_gait_bound)
        c1 = (c[:, f1] > th) & (c[:, f1] < gp.contact_time_gait_bound)
        return torch.where(both | (a0 & c1) | (c0 & a1), 1.0, 0.0)

    def __call__(self, env):
        """rewards.py:116-156"""
        self.update_state(env)
        score = self.task_score(env) if self.gp.encourage_symmetricity_and_low_frequency > 0.5 else None
        sync = (self.sync(env, 0, 1, score) + self.sync(env, 2, 3, score)) / 2.0
        asyn = (self.async_(env, 0, 2) + self.async_(env, 1, 3) + self.async_(env, 0, 3) + self.async_(env, 2, 1)) / 4.0
        stepping = (sync + asyn) / 2.0
        d = env.scene.sensors["robot_contact_senosr"].data
        stance = torch.where(torch.all(d.current_contact_time[:, self.feet] > self.theta, dim=1), 1.0, 0.0)
        stance = stance * self.gp.stance_rwd_scale
        nz = torch.norm(_cmd(env), dim=1) > 0.0
        return torch.where(nz, stepping, stance)


# ------------------------------------------------------------------------------------------------ stateless reward terms
def reward_term(env, kind, p, ids, gait: GaitOracle | None):
    """One raw (unweighted) reward term.  ``ids`` = dict(feet_sensor, feet_body, thigh_calf)."""
    r = env.scene["robot"].data
    cmd = _cmd(env)
    if kind == ALIVE:  # [IL] is_alive
        return (~env.termination_manager.terminated).float()
    if kind == LIN:  # rewards.py:15-20
        return torch.exp(-torch.linalg.norm(cmd[:, :2] - r.root_lin_vel_b[:, :2], dim=1) / p[0])
    if kind == ANG:  # rewards.py:22-27
        return torch.exp(-torch.linalg.norm((cmd[:, 2] - r.root_ang_vel_b[:, 2]).unsqueeze(1), dim=1) / p[0])
    if kind == SLIP:  # rewards.py:31-42
        contact = _contact_mask(env, ids["feet_sensor"], p[0])
        speed = torch.linalg.norm(r.body_lin_vel_w[:, ids["feet_body"], :2], dim=2)
        return torch.sum(contact * speed, dim=1)
    if kind == DRAG:  # rewards.py:44-56
        speed = torch.linalg.norm(r.body_lin_vel_w[:, ids["feet_body"], :2], dim=2)
        low = r.body_pos_w[:, ids["feet_body"], 2] <= p[0]
        return torch.sum(low & (speed > p[1]), dim=1)
    if kind == GAIT:
        return gait(env)
    if kind == HEIGHT:  # rewards.py:398-402
        return torch.square(r.root_pos_w[:, 2] - p[0])
    if kind == ZVEL:  # rewards.py:404-408
        return torch.square(r.root_lin_vel_b[:, 2])
    if kind == RPANG:  # rewards.py:416-420
        return torch.sum(torch.square(r.projected_gravity_b[:, :2]), dim=1)
    if kind == RPVEL:  # rewards.py:410-414
        return torch.sum(torch.abs(r.root_ang_vel_b[:, :2]), dim=1)
    if kind == JLIM:  # rewards.py:423-427
        out = -(r.joint_pos - r.soft_joint_pos_limits[:, :, 0]).clip(max=0.0)
        out += (r.joint_pos - r.soft_joint_pos_limits[:, :, 1]).clip(min=0.0)
        return torch.sum(out, dim=1)
    if kind == JPOS:  # rewards.py:429-440
        c = torch.linalg.norm(cmd, dim=1)
        v = torch.linalg.norm(r.root_lin_vel_b[:, :2], dim=1)
        dev = torch.linalg.norm(r.joint_pos - r.default_joint_pos, dim=1)
        return torch.where((c > 0.0) | (v > p[1]), dev, p[0] * dev)
    if kind == JACC:  # rewards.py:446-448
        return torch.linalg.norm(r.joint_acc, dim=1)
    if kind == JVEL:  # rewards.py:442-444
        return torch.linalg.norm(r.joint_vel, dim=1)
    if kind == JTORQ:  # rewards.py:450-452
        return torch.linalg.norm(r.applied_torque, dim=1)
    if kind == ARATE:  # rewards.py:454-456
        t = env.action_manager.get_term("joint_pos")
        return torch.sum(torch.square(t.raw_actions - t.prev_raw_actions), dim=1)
    if kind == COLL:  # rewards.py:459-466
        return torch.sum(_contact_mask(env, ids["thigh_calf"], p[0]), dim=1)
    # ---- object transport (rewards.py:469-604)
    o = env.scene["object"].data
    q = r.root_quat_w
    nzf = lambda: (torch.linalg.norm(cmd, dim=1) > 0.0)  # noqa: E731
    if kind == OXY:  # :469-481
        dist = torch.linalg.norm((o.root_pos_w - r.root_pos_w)[:, :2], dim=1)
        return dist * nzf() if p and p[0] else dist
    if kind == OXYVEL:  # :483-491
        v = M.quat_apply_inverse(q, o.root_lin_vel_w - r.root_lin_vel_w)
        return torch.sum(torch.square(v[:, :2]), dim=1)
    if kind == OLOSE:  # :596-604
        s = env.scene.sensors["object_contact_sensor"].data
        return torch.logical_and(s.last_contact_time > 0.0, s.current_air_time > 0.0).reshape(-1)
    if kind == OZVEL:  # :493-501
        return torch.square(M.quat_apply_inverse(q, o.root_lin_vel_w - r.root_lin_vel_w)[:, 2])
    if kind in (ORPANG, OROLL):  # :503-512 / :524-533
        g = M.quat_apply_inverse(q, M.quat_apply(o.root_quat_w, o.projected_gravity_b))
        return torch.sum(torch.square(g[:, :2]), dim=1) if kind == ORPANG else torch.square(g[:, 1])
    if kind in (ORPVEL, OROLLVEL):  # :514-522 / :535-543
        w = M.quat_apply_inverse(q, o.root_ang_vel_w - r.root_ang_vel_w)
        return torch.sum(torch.abs(w[:, :2]), dim=1) if kind == ORPVEL else torch.square(w[:, 0])
    if kind == OYAW:  # :545-567
        _, _, ry = M.euler_xyz_from_quat(q)
        _, _, oy = M.euler_xyz_from_quat(o.root_quat_w)
        z = torch.zeros_like(ry)
        qr, qo = M.quat_from_euler_xyz(z, z, ry), M.quat_from_euler_xyz(z, z, oy)
        dyaw = M.euler_xyz_from_quat(M.quat_mul(M.quat_inv(qr), qo))[2]
        dyaw[dyaw > torch.pi] -= 2 * torch.pi
        dyaw[dyaw > 0.5 * torch.pi] -= torch.pi
        dyaw[dyaw <= -0.5 * torch.pi] += torch.pi
        out = torch.square(dyaw)
        return out * nzf() if p and p[0] else out
    if kind == ODANGER:  # :569-594; p = (x_max, y_max, z_min, roll_pitch_max_deg|-1, vel_xy_max|-1)
        rel = M.quat_apply_inverse(q, o.root_pos_w - r.root_pos_w)
        bad = torch.abs(rel[:, 0]) > p[0]
        bad |= torch.abs(rel[:, 1]) > p[1]
        bad |= rel[:, 2] < p[2]
        if p[3] >= 0:
            bad |= torch.acos(-o.projected_gravity_b[:, 2]).abs() > (p[3] * math.pi / 180)
        if p[4] >= 0:
            v = M.quat_apply_inverse(q, o.root_lin_vel_w - r.root_lin_vel_w)
            bad |= torch.linalg.norm(v[:, :2], dim=1) > p[4]
        return bad
    raise ValueError(kind)


def termination_term(env, kind, p, body_ids=None):
    r = env.scene["robot"].data
    if kind == T_TIMEOUT:  # [IL] time_out
        return env.episode_length_buf >= env.max_episode_length
    if kind == T_ORIENT:  # [IL] bad_orientation
        return torch.acos(-r.projected_gravity_b[:, 2]).abs() > p[0]
    if kind == T_HEIGHT:  # [IL] root_height_below_minimum
        return r.root_pos_w[:, 2] < p[0]
    if kind == T_CONTACT:  # [IL] illegal_contact
        return torch.any(_contact_mask(env, body_ids, p[0]), dim=1)
    o = env.scene["object"].data
    if kind == T_BELOW:  # reference terminations.py:10-17
        return o.root_pos_w[:, 2] < r.root_pos_w[:, 2]
    if kind == T_ROLL:  # reference terminations.py:19-23 (on the object)
        return torch.asin(o.projected_gravity_b[:, 1]).abs() > p[0]
    raise ValueError(kind)


def object_state_in_robot_frame(env, os_cfg, noisy, u_state=None, u_euler=None):
    """reference observations.py:38-91.  ``u_state`` [N,13] and ``u_euler`` [N,3] are the explicit uniforms that replace
    ``rand_like`` / ``rand`` (lines 77-78); the never-touched constant gets the *same* additive noise draw row-wise
    (the reference draws a fresh ``rand_like`` at :82 -- a distribution-equivalent but different stream)."""
    r, o = env.scene["robot"].data, env.scene["object"].data
    q = r.root_quat_w
    pos = M.quat_apply_inverse(q, o.root_pos_w - r.root_pos_w)
    lin = M.quat_apply_inverse(q, o.root_lin_vel_w - r.root_lin_vel_w)
    quat = M.quat_mul(M.quat_inv(q), o.root_quat_w)
    ang = M.quat_apply_inverse(q, o.root_ang_vel_w - r.root_ang_vel_w)
    state = torch.cat([pos, lin, quat, ang], dim=-1)
    s = env.scene.sensors["object_contact_sensor"].data
    never = torch.logical_and(s.last_contact_time < os_cfg.last_contact_time_threshold, s.current_contact_time < os_cfg.current_contact_time_threshold).reshape(-1)
    const = torch.tensor(os_cfg.non_contact_obs).repeat(env.num_envs, 1)
    if noisy:
        n_min = torch.tensor(os_cfg.n_min[0:6] + (0.0,) * 4 + os_cfg.n_min[9:])
        n_max = torch.tensor(os_cfg.n_max[0:6] + (0.0,) * 4 + os_cfg.n_max[9:])
        e_min, e_max = torch.tensor(os_cfg.n_min[6:9]), torch.tensor(os_cfg.n_max[6:9])
        add = u_state * (n_max - n_min) + n_min
        state = state + add
        de = u_euler * (e_max - e_min) + e_min
        nq = M.quat_from_euler_xyz(de[:, 0], de[:, 1], de[:, 2])
        state[:, 6:10] = M.quat_mul(state[:, 6:10], nq)
        const = const + add
        const[:, 6:10] = M.quat_mul(const[:, 6:10], nq)
    scale = torch.tensor(os_cfg.scale)
    state = state * scale
    const = const * scale
    return torch.where(never.unsqueeze(-1), const, state)


def action_term_process(term, actions, clip, raw_scale, scale, offset):
    """JointPositionActionPrevPrev.process_actions (reference locotouch/mdp/actions.py:30-44) over [IL] JointAction.process_actions
    (raw[:] = actions; processed = raw * scale + offset): the histories shift first, then clamp (``clip`` None: off), raw scale."""
    term.prev_prev_raw_actions[:] = term.prev_raw_actions.clone()
    term.prev_raw_actions[:] = term.raw_actions.clone()
    if getattr(term, "prev_processed_actions", None) is not None:
        term.prev_prev_processed_actions[:] = term.prev_processed_actions.clone()
        term.prev_processed_actions[:] = term.processed_actions.clone()
    if clip is not None:
        actions = torch.clamp(actions, -clip, clip)
    actions = actions * raw_scale
    term.raw_actions[:] = actions
    term.processed_actions[:] = term.raw_actions * scale + offset


def action_term_reset(term, env_ids, scale=None, offset=None):
    """JointPositionActionPrevPrev.reset (reference locotouch/mdp/actions.py:46-52) on the state block the hot path reads:
    prev / prev_prev raw actions zeroed, processed = raw * scale + offset of the PRE-reset raw action (line 49 runs before the base
    class), then [IL] ActionTerm.reset zeroes raw_actions of those envs.  ``scale`` / ``offset`` None: processed_actions is left as
    process_actions wrote it (the same value by construction)."""
    term.prev_raw_actions[env_ids] = 0.0
    if getattr(term, "prev_prev_raw_actions", None) is not None:
        term.prev_prev_raw_actions[env_ids] = 0.0
    if scale is not None and getattr(term, "processed_actions", None) is not None:
        term.processed_actions[env_ids] = (term.raw_actions * scale + (0.0 if offset is None else offset))[env_ids]
    if getattr(term, "prev_processed_actions", None) is not None:
        term.prev_processed_actions[env_ids] = term.processed_actions[env_ids].clone()
        term.prev_prev_processed_actions[env_ids] = term.processed_actions[env_ids].clone()
    term.raw_actions[env_ids] = 0.0


class MdpOracle:
    """Managers + terms for one task: ``step(env)`` = terminations -> rewards -> (auto reset) ; ``observe(env)``."""

    def __init__(self, env, spec):
        self.spec = spec
        n = env.num_envs
        sensor = env.scene.sensors["robot_contact_senosr"]
        robot = env.scene["robot"]
        pair0 = sensor.find_bodies(list(spec.gait.synced_feet_pair_names[0]))[0]
        pair1 = sensor.find_bodies(list(spec.gait.synced_feet_pair_names[1]))[0]
        self.gait = GaitOracle(env, spec.gait, [pair0[0], pair0[1], pair1[0], pair1[1]])
        self.ids = dict(
            feet_sensor=sensor.find_bodies(".*foot")[0],
            feet_body=robot.find_bodies(".*foot")[0],
            thigh_calf=sensor.find_bodies(list(spec.thigh_calf_names))[0],
        )
        self.term_body_ids = {t.name: (sensor.find_bodies(t.body_names)[0] if t.body_names else None) for t in spec.terminations}
        nt = len(spec.rewards)
        self.episode_sums = torch.zeros(nt, n)
        self.step_reward = torch.zeros(n, nt)
        self.reward_buf = torch.zeros(n)
        d = spec.obs_dim
        self.policy_obs = torch.zeros(n, d)
        self.critic_obs = torch.zeros(n, d)
        self.needs_fill = torch.ones(n, dtype=torch.bool)  # history ring empty -> first push fills all slots [IL]

    # -- [IL] TerminationManager.compute
    def terminations(self, env):
        n = env.num_envs
        masks = {}
        time_outs = torch.zeros(n, dtype=torch.bool)
        terminated = torch.zeros(n, dtype=torch.bool)
        for t in self.spec.terminations:
            m = termination_term(env, t.kind, t.p, self.term_body_ids[t.name])
            masks[t.name] = m
            if t.time_out:
                time_outs |= m
            else:
                terminated |= m
        env.termination_manager.terminated = terminated
        env.termination_manager.time_outs = time_outs
        return masks, terminated, time_outs

    # -- [IL] RewardManager.compute(dt): sequential fp32 accumulation in cfg order, zero-weight terms skipped
    def rewards(self, env):
        dt = env.step_dt
        self.reward_buf[:] = 0.0
        raw = {}
        for i, t in enumerate(self.spec.rewards):
            if t.weight == 0.0:
                self.step_reward[:, i] = 0.0
                continue
            val = reward_term(env, t.kind, t.p, self.ids, self.gait)
            raw[t.name] = val
            value = val * t.weight * dt
            self.reward_buf += value
            self.episode_sums[i] += value
            self.step_reward[:, i] = value / dt
        return raw, self.reward_buf

    def step(self, env, auto_reset=True, reset_action_term=False):
        """``reset_action_term``: [IL] ManagerBasedRLEnv.step also runs ActionManager.reset(env_ids) between the reward and the
        observation pass; for the LocoTouch action term that is JointPositionActionPrevPrev.reset (reference locotouch/mdp/actions.py:46-52)
        followed by [IL] ActionTerm.reset (raw_actions[env_ids] = 0)."""
        masks, terminated, time_outs = self.terminations(env)
        raw, reward = self.rewards(env)
        reward = reward.clone()
        done = terminated | time_outs
        if auto_reset:
            ids = done.nonzero(as_tuple=False).flatten()
            if len(ids) > 0:
                self.gait.reset(ids)
                self.episode_sums[:, ids] = 0.0
                self.needs_fill[ids] = True
                if reset_action_term:
                    action_term_reset(env.action_manager.get_term("joint_pos"), ids)
        return dict(masks=masks, terminated=terminated, time_outs=time_outs, raw=raw, reward=reward, done=done)

    # -- [IL] ObservationManager.compute for the policy (noisy) and critic (clean) groups
    def observe(self, env, u_noise=None, u_obj_euler=None):
        """``u_noise`` [N, obs_dim_per_step] uniforms in [0,1) for the additive noise of the policy group."""
        r = env.scene["robot"].data
        term = env.action_manager.get_term("joint_pos")
        n = env.num_envs
        raw_terms = {
            "velocity_commands": _cmd(env),
            "base_ang_vel": r.root_ang_vel_b,
            "projected_gravity": r.projected_gravity_b,
            "joint_pos": r.joint_pos - r.default_joint_pos,
            "joint_vel": r.joint_vel - r.default_joint_vel,
            "last_action": term.raw_actions,
        }
        hl = self.spec.history_length
        out = []
        for noisy, buf in ((True, self.policy_obs), (False, self.critic_obs)):
            col, ucol = 0, 0
            new_buf = torch.empty_like(buf)
            for t in self.spec.obs_terms:
                if t.name == "object_state":
                    u_state = u_noise[:, ucol : ucol + t.dim] if noisy else None  # the term's own 13 uniform columns
                    v = object_state_in_robot_frame(env, self.spec.object_state, noisy, u_state, u_obj_euler)
                else:
                    v = raw_terms[t.name].clone()
                    if noisy and t.noise is not None:
                        lo, hi = t.noise
                        v = v + u_noise[:, ucol : ucol + t.dim] * (hi - lo) + lo  # [IL] data + rand*(n_max-n_min) + n_min
                    v = v * t.scale
                ucol += t.dim
                old = buf[:, col : col + hl * t.dim].view(n, hl, t.dim)
                shifted = torch.cat([old[:, 1:], v.unsqueeze(1)], dim=1)
                filled = v.unsqueeze(1).expand(n, hl, t.dim)
                res = torch.where(self.needs_fill.view(n, 1, 1), filled, shifted)
                new_buf[:, col : col + hl * t.dim] = res.reshape(n, hl * t.dim)
                col += hl * t.dim
            buf.copy_(new_buf)
            out.append(buf.clone())
        self.needs_fill[:] = False
        return out[0], out[1]
