"""Seeded synthetic Go1 / LocoTouch state (stand-in for PhysX output) -- BASELINE.json configs C2, C3, C4.

Distributions follow SURVEY.md section 8(d).  Everything is drawn on the CPU from an explicit
``torch.Generator`` and then moved, so that the CPU oracle and the CUDA path read identical bits.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch

from .scene import (
    ROBOT_BODY_NAMES,
    ROBOT_JOINT_NAMES,
    TAXEL_BODY_NAMES,
    TAXEL_COLS,
    TAXEL_ROWS,
    ActionTermState,
    Entity,
    SynthEnv,
)

NUM_JOINTS = 12
NUM_ROBOT_BODIES = 17
FORCE_HISTORY = 3  # reference locomotion_base_env_cfg.py:37
# Go1 default joint positions (reference assets/go1.py:32-37), joint order = ROBOT_JOINT_NAMES
DEFAULT_JOINT_POS = [-0.1, 0.1, -0.1, 0.1] + [0.9] * 4 + [-1.8] * 4
# Go1 URDF joint ranges (hip, thigh, calf) scaled by soft_joint_pos_limit_factor=0.95 (reference assets/go1.py:29)
_URDF_LIMITS = [(-0.863, 0.863)] * 4 + [(-0.686, 4.501)] * 4 + [(-2.818, -0.888)] * 4


def _soft_limits() -> torch.Tensor:
    lim = torch.tensor(_URDF_LIMITS, dtype=torch.float32)
    mean = lim.mean(dim=1)
    half = 0.5 * (lim[:, 1] - lim[:, 0]) * 0.95
    return torch.stack([mean - half, mean + half], dim=1)


def _randn(gen, *shape, std=1.0, mean=0.0):
    return torch.randn(*shape, generator=gen, dtype=torch.float32) * std + mean


def _rand(gen, *shape, lo=0.0, hi=1.0):
    return torch.rand(*shape, generator=gen, dtype=torch.float32) * (hi - lo) + lo


def _quat_from_euler(roll, pitch, yaw):
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    return torch.stack(
        [cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp],
        dim=-1,
    )


def _rotate(q, v):
    w, xyz = q[..., :1], q[..., 1:]
    t = 2.0 * torch.cross(xyz, v, dim=-1)
    return v + w * t + torch.cross(xyz, t, dim=-1)


def draw_robot_state(env: SynthEnv, gen: torch.Generator, zero_cmd_frac: float = 0.1, keep_cmd: bool = False):
    """(Re)draw the kinematic state of the robot: one synthetic 'PhysX step' worth of tensors (config C2)."""
    n = env.num_envs
    robot = env.scene["robot"].data
    if not keep_cmd:
        cmd = torch.stack(
            [_rand(gen, n, lo=-1.0, hi=1.0), _rand(gen, n, lo=-0.6, hi=0.6), _rand(gen, n, lo=-math.pi / 2, hi=math.pi / 2)], dim=1
        )
        cmd[torch.rand(n, generator=gen) < zero_cmd_frac] = 0.0
        env.command_manager.set_command("base_velocity", cmd)
    robot.root_pos_w = torch.cat([_randn(gen, n, 2, std=2.0), _randn(gen, n, 1, std=0.03, mean=0.30)], dim=1)
    robot.root_quat_w = _quat_from_euler(_randn(gen, n, std=0.1), _randn(gen, n, std=0.1), _rand(gen, n, lo=-math.pi, hi=math.pi))
    robot.root_lin_vel_b = _randn(gen, n, 3, std=0.5)
    robot.root_ang_vel_b = _randn(gen, n, 3, std=0.5)
    robot.root_lin_vel_w = _rotate(robot.root_quat_w, robot.root_lin_vel_b)
    robot.root_ang_vel_w = _rotate(robot.root_quat_w, robot.root_ang_vel_b)
    g = torch.cat([_randn(gen, n, 2, std=0.1), -torch.ones(n, 1)], dim=1)
    # a few tipped-over robots so that bad_orientation fires
    tipped = torch.rand(n, generator=gen) < 0.003
    g[tipped, 2] = _rand(gen, int(tipped.sum()), lo=0.0, hi=0.5)
    robot.projected_gravity_b = g / g.norm(dim=1, keepdim=True)
    q0 = torch.tensor(DEFAULT_JOINT_POS).repeat(n, 1)
    robot.default_joint_pos = q0
    robot.default_joint_vel = torch.zeros(n, NUM_JOINTS)
    robot.joint_pos = q0 + _rand(gen, n, NUM_JOINTS, lo=-0.3, hi=0.3)
    # push a few joints beyond the soft limits so joint_position_limit is non-zero
    robot.joint_pos += (torch.rand(n, NUM_JOINTS, generator=gen) < 0.02) * _randn(gen, n, NUM_JOINTS, std=1.0)
    robot.joint_vel = _randn(gen, n, NUM_JOINTS, std=2.0)
    robot.joint_acc = _randn(gen, n, NUM_JOINTS, std=50.0)
    robot.applied_torque = _rand(gen, n, NUM_JOINTS, lo=-23.5, hi=23.5)
    robot.soft_joint_pos_limits = _soft_limits().repeat(n, 1, 1)
    nb = env.scene["robot"].num_bodies
    body_pos = robot.root_pos_w.unsqueeze(1) + _randn(gen, n, nb, 3, std=0.2)
    feet = [ROBOT_BODY_NAMES.index(f"{leg}_foot") for leg in ("a_FR", "b_FL", "c_RR", "d_RL")]
    body_pos[:, feet, 2] = _rand(gen, n, 4, lo=0.0, hi=0.1)
    robot.body_pos_w = body_pos
    robot.body_lin_vel_w = _randn(gen, n, nb, 3, std=0.5)
    low_speed = torch.rand(n, nb, generator=gen) < 0.3
    robot.body_lin_vel_w[low_speed] *= 0.05  # some near-stationary feet (exercise the 0.1 m/s dragging threshold)
    root_z_low = torch.rand(n, generator=gen) < 0.003
    robot.root_pos_w[root_z_low, 2] = _rand(gen, int(root_z_low.sum()), lo=0.05, hi=0.15)
    # episode counters, some of them at the time-out boundary
    env.episode_length_buf = torch.randint(0, env.max_episode_length, (n,), generator=gen, dtype=torch.long)
    at_limit = torch.rand(n, generator=gen) < 0.003
    env.episode_length_buf[at_limit] = env.max_episode_length


def draw_actions(env: SynthEnv, gen: torch.Generator):
    """JointPositionActionPrevPrev bookkeeping for a new policy output (reference mdp/actions.py:30-44)."""
    term: ActionTermState = env.action_manager.get_term("joint_pos")
    n = env.num_envs
    policy_out = _randn(gen, n, NUM_JOINTS)
    term.prev_prev_raw_actions = term.prev_raw_actions.clone()
    term.prev_raw_actions = term.raw_actions.clone()
    term.raw_actions = torch.clamp(policy_out, -100.0, 100.0) * 0.25
    term.processed_actions = term.raw_actions + env.scene["robot"].data.default_joint_pos
    return policy_out


def _contact_prob(sensor) -> torch.Tensor:
    """Per-body probability of being in contact: feet often, thighs/calves sometimes, hips/trunk rarely."""
    p = []
    for name in sensor.body_names:
        if name.endswith("foot"):
            p.append(0.5)
        elif name.endswith("thigh") or name.endswith("calf"):
            p.append(0.04)
        else:
            p.append(0.0008)
    return torch.tensor(p)


def init_contacts(env: SynthEnv, gen: torch.Generator, sensor_name: str = "robot_contact_senosr"):
    """ContactSensor air/contact-time state: mutually exclusive accumulators, latched ``last_*`` ([IL])."""
    n = env.num_envs
    sensor = env.scene.sensors[sensor_name]
    nb = sensor.num_bodies
    d = sensor.data
    in_contact = torch.rand(n, nb, generator=gen) < _contact_prob(sensor)
    t = _rand(gen, n, nb, lo=0.0, hi=0.6)
    d.current_contact_time = torch.where(in_contact, t, torch.zeros_like(t))
    d.current_air_time = torch.where(in_contact, torch.zeros_like(t), t)
    d.last_air_time = _rand(gen, n, nb, lo=0.0, hi=0.6)
    d.last_contact_time = _rand(gen, n, nb, lo=0.0, hi=0.6)
    _draw_forces(env, gen, sensor_name, first=True)


def _draw_forces(env, gen, sensor_name, first=False):
    n = env.num_envs
    sensor = env.scene.sensors[sensor_name]
    nb = sensor.num_bodies
    d = sensor.data
    in_contact = d.current_contact_time > 0.0
    frame = _randn(gen, n, nb, 3, std=30.0) * in_contact.unsqueeze(-1)
    # occasional grazing contacts around the 0.1 / 0.5 / 1.0 N thresholds
    grazing = torch.rand(n, nb, generator=gen) < _contact_prob(sensor) * 0.2
    frame = torch.where(grazing.unsqueeze(-1), _randn(gen, n, nb, 3, std=0.4), frame)
    if first or not hasattr(d, "net_forces_w_history"):
        hist = torch.stack([frame] + [_randn(gen, n, nb, 3, std=30.0) * in_contact.unsqueeze(-1) for _ in range(FORCE_HISTORY - 1)], dim=1)
    else:
        hist = torch.cat([frame.unsqueeze(1), d.net_forces_w_history[:, :-1]], dim=1)
    d.net_forces_w_history = hist.contiguous()
    d.net_forces_w = frame.contiguous()


def advance_contacts(env: SynthEnv, gen: torch.Generator, toggle_prob: float = 0.2, sensor_name: str = "robot_contact_senosr"):
    """One env step of the contact state machine: accumulate dt, toggle some bodies, latch ``last_*`` on transition."""
    d = env.scene.sensors[sensor_name].data
    dt = env.step_dt
    n, nb = d.current_air_time.shape
    in_contact = d.current_contact_time > 0.0
    p_contact = _contact_prob(env.scene.sensors[sensor_name])
    # stationary contact fraction stays at p_contact: P(land) = toggle * p/(1-p) scaled, P(lift) = toggle
    p_land = (toggle_prob * p_contact / (1.0 - p_contact)).clamp(max=1.0)
    u = torch.rand(n, nb, generator=gen)
    toggle = torch.where(in_contact, u < toggle_prob, u < p_land)
    landing = toggle & ~in_contact
    lifting = toggle & in_contact
    d.last_air_time = torch.where(landing, d.current_air_time + dt, d.last_air_time)
    d.last_contact_time = torch.where(lifting, d.current_contact_time + dt, d.last_contact_time)
    new_contact = (in_contact & ~lifting) | landing
    cct = torch.where(landing, torch.full_like(d.current_contact_time, dt), d.current_contact_time + dt)
    cat = torch.where(lifting, torch.full_like(d.current_air_time, dt), d.current_air_time + dt)
    d.current_contact_time = torch.where(new_contact, cct, torch.zeros_like(cct))
    d.current_air_time = torch.where(new_contact, torch.zeros_like(cat), cat)
    _draw_forces(env, gen, sensor_name)


def draw_object_state(env: SynthEnv, gen: torch.Generator, never_touched_frac: float = 0.05):
    """Object on the robot's back (config C3): pose relative to the trunk, relative velocities, contact times."""
    n = env.num_envs
    robot = env.scene["robot"].data
    obj = env.scene["object"].data
    rel = torch.stack([_rand(gen, n, lo=-0.15, hi=0.15), _rand(gen, n, lo=-0.12, hi=0.12), _rand(gen, n, lo=0.08, hi=0.2)], dim=1)
    obj.root_pos_w = robot.root_pos_w + _rotate(robot.root_quat_w, rel)
    fallen = torch.rand(n, generator=gen) < 0.003
    obj.root_pos_w[fallen, 2] = robot.root_pos_w[fallen, 2] - _rand(gen, int(fallen.sum()), lo=0.01, hi=0.2)
    roll = _randn(gen, n, std=0.3)
    rolled = torch.rand(n, generator=gen) < 0.003
    roll[rolled] = _rand(gen, int(rolled.sum()), lo=1.1, hi=1.5)
    obj.root_quat_w = _quat_from_euler(roll, _rand(gen, n, lo=-math.pi, hi=math.pi), _rand(gen, n, lo=-math.pi, hi=math.pi))
    obj.root_lin_vel_w = robot.root_lin_vel_w + _randn(gen, n, 3, std=0.3)
    fast = torch.rand(n, generator=gen) < 0.01
    obj.root_lin_vel_w[fast] += _randn(gen, int(fast.sum()), 3, std=3.0)
    obj.root_ang_vel_w = _randn(gen, n, 3, std=1.0)
    # gravity direction in the object frame = R(q_o)^T (0,0,-1)
    qc = obj.root_quat_w * torch.tensor([1.0, -1.0, -1.0, -1.0])
    obj.projected_gravity_b = _rotate(qc, torch.tensor([0.0, 0.0, -1.0]).repeat(n, 1))
    s = env.scene.sensors["object_contact_sensor"].data
    touching = torch.rand(n, 1, generator=gen) < 0.8
    t = _rand(gen, n, 1, lo=0.0, hi=2.0)
    s.current_contact_time = torch.where(touching, t, torch.zeros_like(t))
    s.current_air_time = torch.where(touching, torch.zeros_like(t), _rand(gen, n, 1, lo=0.0, hi=0.3))
    s.last_contact_time = _rand(gen, n, 1, lo=0.0, hi=2.0)
    s.last_air_time = _rand(gen, n, 1, lo=0.0, hi=0.3)
    never = torch.rand(n, 1, generator=gen) < never_touched_frac
    s.current_contact_time[never] = 0.0
    s.last_contact_time[never] = 0.0


def draw_tactile_state(env: SynthEnv, gen: torch.Generator, jitter: float = 0.0):
    """221 taxel links (config C4): a cylinder-shaped footprint pressing on the back pad.

    Sensor link quaternions equal the trunk quaternion (fixed joints, identity offset -- reference
    generate_locotouch_urdf.py:59-72); ``jitter`` adds a per-taxel rotation to exercise the general path.
    """
    n = env.num_envs
    robot = env.scene["robot"]
    nt = TAXEL_ROWS * TAXEL_COLS
    q_trunk = robot.data.root_quat_w
    body_quat = q_trunk.unsqueeze(1).repeat(1, nt, 1)
    if jitter > 0.0:
        dq = _quat_from_euler(_randn(gen, n, nt, std=jitter), _randn(gen, n, nt, std=jitter), _randn(gen, n, nt, std=jitter))
        w1, x1, y1, z1 = body_quat.unbind(-1)
        w2, x2, y2, z2 = dq.unbind(-1)
        body_quat = torch.stack(
            [w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2],
            dim=-1,
        )
    # all articulation bodies: the 17 structural links followed by the taxel links
    n_struct = NUM_ROBOT_BODIES
    full = torch.zeros(n, n_struct + nt, 4)
    full[:, :n_struct] = q_trunk.unsqueeze(1)
    full[:, n_struct:] = body_quat
    robot.data.body_quat_w = full.contiguous()
    # footprint: a line of ~3 x 10 taxels at a random pose
    rr = torch.arange(TAXEL_ROWS, dtype=torch.float32).view(1, -1, 1)
    cc = torch.arange(TAXEL_COLS, dtype=torch.float32).view(1, 1, -1)
    cx, cy = _rand(gen, n, 1, 1, lo=3.0, hi=13.0), _rand(gen, n, 1, 1, lo=3.0, hi=9.0)
    th = _rand(gen, n, 1, 1, lo=-0.6, hi=0.6)
    du, dv = rr - cx, cc - cy
    along = du * torch.cos(th) + dv * torch.sin(th)
    across = -du * torch.sin(th) + dv * torch.cos(th)
    inside = (along.abs() <= 5.0) & (across.abs() <= 1.5)
    f_local = torch.zeros(n, TAXEL_ROWS, TAXEL_COLS, 3)
    f_local[..., 2] = -_rand(gen, n, TAXEL_ROWS, TAXEL_COLS, lo=0.0, hi=3.0) * inside
    # weak presses straddling the 0.05 N contact threshold
    weak = torch.rand(n, TAXEL_ROWS, TAXEL_COLS, generator=gen) < 0.05
    f_local[..., 2] = torch.where(weak, -_rand(gen, n, TAXEL_ROWS, TAXEL_COLS, lo=0.03, hi=0.07), f_local[..., 2])
    f_local = f_local.reshape(n, nt, 3) + _randn(gen, n, nt, 3, std=0.01)
    tac = env.scene.sensors["tactile_contact_sensor"].data
    tac.net_forces_w = _rotate(body_quat, f_local).contiguous()


def default_reward_term_cfgs(with_object: bool) -> dict[str, SimpleNamespace]:
    """``reward_manager.get_term_cfg`` content the gait-with-object term looks up (reference rewards.py:380-381)."""
    cfgs = {}
    if with_object:
        cfgs["object_dangerous_state"] = SimpleNamespace(
            params={"x_max": 0.125, "y_max": 0.097, "z_min": 0.095, "roll_pitch_max": None, "vel_xy_max": 2.5}, weight=-50.0
        )
    return cfgs


def make_env(
    num_envs: int,
    seed: int = 0,
    with_object: bool = False,
    with_tactile: bool = False,
    tactile_jitter: float = 0.0,
    max_episode_length: int | None = None,
) -> SynthEnv:
    """CPU SynthEnv for configs C2 (locomotion), C3 (``with_object``) and C4 (``with_tactile``)."""
    if max_episode_length is None:
        max_episode_length = 500 if with_tactile else 1000  # reference object_transport_student_env_cfg.py:176
    env = SynthEnv(num_envs, "cpu", step_dt=0.02, max_episode_length=max_episode_length)
    gen = torch.Generator().manual_seed(seed)
    env.generator = gen
    body_names = ROBOT_BODY_NAMES + (TAXEL_BODY_NAMES if with_tactile else [])
    env.scene["robot"] = Entity(body_names, ROBOT_JOINT_NAMES)
    env.scene.sensors["robot_contact_senosr"] = Entity(ROBOT_BODY_NAMES)
    env.action_manager._terms["joint_pos"] = ActionTermState(num_envs, NUM_JOINTS)
    draw_robot_state(env, gen)
    draw_actions(env, gen)
    draw_actions(env, gen)
    init_contacts(env, gen)
    if with_object:
        env.scene["object"] = Entity(["Object"])
        env.scene.sensors["object_contact_sensor"] = Entity(["Object"])
        draw_object_state(env, gen)
    if with_tactile:
        env.scene.sensors["tactile_contact_sensor"] = Entity(TAXEL_BODY_NAMES)
        draw_tactile_state(env, gen, tactile_jitter)
    for name, cfg in default_reward_term_cfgs(with_object).items():
        env.reward_manager.set_term_cfg(name, cfg)
    return env


def advance(env: SynthEnv, keep_cmd_prob: float = 0.8, tactile_jitter: float = 0.0):
    """Synthetic replacement of one ``action -> PhysX x4`` transition: new kinematic state, evolved contacts."""
    gen = env.generator
    keep = bool(torch.rand((), generator=gen) < keep_cmd_prob)
    # [IL] the env zeroes episode_length_buf of the envs it reset at the end of the previous step
    prev_len = torch.where(env.termination_manager.dones.cpu(), 0, env.episode_length_buf)
    draw_robot_state(env, gen, keep_cmd=keep)
    env.episode_length_buf = prev_len + 1
    draw_actions(env, gen)
    advance_contacts(env, gen)
    if "object" in env.scene._entities:
        draw_object_state(env, gen)
    if "tactile_contact_sensor" in env.scene.sensors:
        draw_tactile_state(env, gen, tactile_jitter)
    env.common_step_counter += 1
