"""Declarative spec of the MDP terms of the two LocoTouch tasks on the hot path.

This is the *parameter* content of the reference's config trees (which are IsaacLab ``@configclass`` objects and
therefore not importable without the simulator), restated as plain data:

* ``Isaac-Locomotion-LocoTouch-v1``           reference locotouch/config/base/locomotion_base_env_cfg.py:139-218
  (17 reward terms), :296-313 (5 terminations), :70-122 (6 observation terms x history 6 -> 270).
* ``Isaac-RandCylinderTransportTeacher-LocoTouch-v1``  adds the object terms of
  reference locotouch/config/locotouch/object_transport_teacher_env_cfg.py:88-114 with the cylinder overrides of
  cylinder_transport_teacher_env_cfg.py:41-52 (23 active rewards, 6 terminations, 348-D observations).

The integer ``kind`` codes are shared with the CUDA kernel (csrc/mdp_step.cu, ``enum LtRewardKind``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

# ---------------------------------------------------------------- reward kinds (keep in sync with include/locotouch_b200.h)
RK_ALIVE = 0
RK_TRACK_LIN_VEL_XY = 1
RK_TRACK_ANG_VEL_Z = 2
RK_FOOT_SLIP = 3
RK_FOOT_DRAG = 4
RK_GAIT = 5
RK_BASE_HEIGHT = 6
RK_BASE_Z_VEL = 7
RK_BASE_RP_ANGLE = 8
RK_BASE_RP_VEL = 9
RK_JOINT_POS_LIMIT = 10
RK_JOINT_POS = 11
RK_JOINT_ACC = 12
RK_JOINT_VEL = 13
RK_JOINT_TORQUE = 14
RK_ACTION_RATE = 15
RK_THIGH_CALF_COLLISION = 16
RK_OBJ_XY_POS = 17
RK_OBJ_XY_VEL = 18
RK_OBJ_LOSE_CONTACT = 19
RK_OBJ_Z_VEL = 20
RK_OBJ_RP_ANGLE = 21
RK_OBJ_RP_VEL = 22
RK_OBJ_ROLL_ANGLE = 23
RK_OBJ_ROLL_VEL = 24
RK_OBJ_YAW = 25
RK_OBJ_DANGER = 26

# ---------------------------------------------------------------- termination kinds
TK_TIME_OUT = 0
TK_BAD_ORIENTATION = 1
TK_ROOT_HEIGHT = 2
TK_ILLEGAL_CONTACT = 3
TK_OBJECT_BELOW_ROBOT = 4
TK_BAD_ROLL = 5


@dataclass
class RewardTerm:
    name: str
    kind: int
    weight: float
    p: tuple = ()  # up to 6 float parameters, meaning depends on kind


@dataclass
class TerminationTerm:
    name: str
    kind: int
    time_out: bool = False
    p: tuple = ()
    body_names: object = None  # for illegal_contact


@dataclass
class ObsTerm:
    name: str
    dim: int
    scale: float = 1.0
    noise: tuple | None = None  # (n_min, n_max) additive uniform, applied before scale ([IL] order noise->clip->scale)


@dataclass
class GaitParams:
    """reference locomotion_base_env_cfg.py:168-188"""

    synced_feet_pair_names: tuple = (("a_FR_foot", "d_RL_foot"), ("b_FL_foot", "c_RR_foot"))
    judge_time_threshold: float = 1.0e-6
    air_time_gait_bound: float = 0.5
    contact_time_gait_bound: float = 0.5
    async_time_tolerance: float = 0.05
    stance_rwd_scale: float = 1.0
    encourage_symmetricity_and_low_frequency: float = 1.0
    soft_minimum_frequency: float = 2.0
    tolerance_proportion: float = 0.2
    rwd_upper_bound: float = 1.0
    rwd_lower_bound: float = -5.0
    vel_tracking_exp_sigma: float = 0.25
    task_performance_ratio: float = 1.0
    with_object: bool = False
    obj_x_max: float = 0.125  # read from the object_dangerous_state term cfg (reference rewards.py:380-381)
    obj_y_max: float = 0.097


@dataclass
class ObjectStateObs:
    """reference object_transport_teacher_env_cfg.py:14-30"""

    n_min: tuple = (-0.01, -0.01, -0.005) + (-0.2,) * 3 + (-0.05,) * 3 + (-0.2,) * 3
    n_max: tuple = (0.01, 0.01, 0.005) + (0.2,) * 3 + (0.05,) * 3 + (0.2,) * 3
    scale: tuple = (1.0,) * 3 + (0.5,) * 3 + (1.0,) * 4 + (0.25,) * 3
    non_contact_obs: tuple = (0.0,) * 6 + (1.0,) + (0.0,) * 6
    last_contact_time_threshold: float = 1e-8
    current_contact_time_threshold: float = 1e-8


@dataclass
class TaskSpec:
    name: str
    rewards: list[RewardTerm]
    terminations: list[TerminationTerm]
    obs_terms: list[ObsTerm]
    gait: GaitParams
    object_state: ObjectStateObs | None = None
    history_length: int = 6
    step_dt: float = 0.02
    max_episode_length: int = 1000
    feet_names: tuple = ("a_FR_foot", "b_FL_foot", "c_RR_foot", "d_RL_foot")  # regex ".*foot" in body order
    thigh_calf_names: tuple = (".*thigh", ".*calf")

    @property
    def with_object(self) -> bool:
        return self.object_state is not None

    @property
    def obs_dim_per_step(self) -> int:
        return sum(t.dim for t in self.obs_terms)

    @property
    def obs_dim(self) -> int:
        return self.obs_dim_per_step * self.history_length

    @property
    def active_rewards(self) -> list[RewardTerm]:
        """[IL] RewardManager.compute skips terms whose weight is exactly 0.0."""
        return [t for t in self.rewards if t.weight != 0.0]


def _base_rewards() -> list[RewardTerm]:
    return [
        RewardTerm("alive", RK_ALIVE, 10.0),
        RewardTerm("track_lin_vel_xy", RK_TRACK_LIN_VEL_XY, 1.0, (0.25,)),
        RewardTerm("track_ang_vel_z", RK_TRACK_ANG_VEL_Z, 0.5, (0.25,)),
        RewardTerm("foot_slip", RK_FOOT_SLIP, -1.0, (0.5,)),
        RewardTerm("foot_dragging", RK_FOOT_DRAG, -0.1, (0.03, 0.1)),
        RewardTerm("gait", RK_GAIT, 0.5),
        RewardTerm("track_base_height", RK_BASE_HEIGHT, -0.5, (0.42,)),
        RewardTerm("base_z_velocity", RK_BASE_Z_VEL, -1.0),
        RewardTerm("base_roll_pitch_angle", RK_BASE_RP_ANGLE, -1.0),
        RewardTerm("base_roll_pitch_velocity", RK_BASE_RP_VEL, -0.2),
        RewardTerm("joint_position_limit", RK_JOINT_POS_LIMIT, -10.0),
        RewardTerm("joint_position", RK_JOINT_POS, -0.5, (5.0, 0.3)),
        RewardTerm("joint_acceleration", RK_JOINT_ACC, -5.0e-6),
        RewardTerm("joint_velocity", RK_JOINT_VEL, -5.0e-3),
        RewardTerm("joint_torque", RK_JOINT_TORQUE, -2.5e-4),
        RewardTerm("action_rate", RK_ACTION_RATE, -0.75),
        RewardTerm("thigh_calf_collision", RK_THIGH_CALF_COLLISION, -5.0, (0.1,)),
    ]


def _base_obs() -> list[ObsTerm]:
    return [
        ObsTerm("velocity_commands", 3, 1.0, None),
        ObsTerm("base_ang_vel", 3, 0.25, (-0.2, 0.2)),
        ObsTerm("projected_gravity", 3, 1.0, (-0.05, 0.05)),
        ObsTerm("joint_pos", 12, 1.0, (-0.01, 0.01)),
        ObsTerm("joint_vel", 12, 0.05, (-1.5, 1.5)),
        ObsTerm("last_action", 12, 1.0, None),
    ]


def locomotion_spec() -> TaskSpec:
    """Isaac-Locomotion-LocoTouch-v1"""
    return TaskSpec(
        name="Isaac-Locomotion-LocoTouch-v1",
        rewards=_base_rewards(),
        terminations=[
            TerminationTerm("time_out", TK_TIME_OUT, time_out=True),
            TerminationTerm("base_orientation", TK_BAD_ORIENTATION, p=(math.pi / 2,)),
            TerminationTerm("base_height_below_minimum", TK_ROOT_HEIGHT, p=(0.15,)),
            TerminationTerm("base_contact", TK_ILLEGAL_CONTACT, p=(1.0,), body_names="trunk"),
            TerminationTerm("hip_contact", TK_ILLEGAL_CONTACT, p=(1.0,), body_names=".*hip"),
        ],
        obs_terms=_base_obs(),
        gait=GaitParams(),
    )


def teacher_spec() -> TaskSpec:
    """Isaac-RandCylinderTransportTeacher-LocoTouch-v1"""
    rewards = _base_rewards()
    rewards += [
        RewardTerm("object_xy_position", RK_OBJ_XY_POS, -50.0, (1.0,)),
        RewardTerm("object_xy_velocity", RK_OBJ_XY_VEL, 0.0),
        RewardTerm("object_z_contact", RK_OBJ_LOSE_CONTACT, 0.0),
        RewardTerm("object_z_velocity", RK_OBJ_Z_VEL, -0.5),
        RewardTerm("object_roll_pitch_angle", RK_OBJ_ROLL_ANGLE, -0.05),
        RewardTerm("object_roll_pitch_velocity", RK_OBJ_ROLL_VEL, -0.05),
        RewardTerm("object_yaw_alignment", RK_OBJ_YAW, -0.1, (1.0,)),
        # p = (x_max, y_max, z_min, roll_pitch_max_deg or -1 for None, vel_xy_max)
        RewardTerm("object_dangerous_state", RK_OBJ_DANGER, -50.0, (0.125, 0.097, 0.095, -1.0, 2.5)),
    ]
    return TaskSpec(
        name="Isaac-RandCylinderTransportTeacher-LocoTouch-v1",
        rewards=rewards,
        terminations=[
            TerminationTerm("time_out", TK_TIME_OUT, time_out=True),
            TerminationTerm("base_orientation", TK_BAD_ORIENTATION, p=(math.pi / 2,)),
            TerminationTerm("base_height_below_minimum", TK_ROOT_HEIGHT, p=(0.15,)),
            TerminationTerm("hip_contact", TK_ILLEGAL_CONTACT, p=(1.0,), body_names=".*hip"),
            TerminationTerm("object_below_robot", TK_OBJECT_BELOW_ROBOT),
            TerminationTerm("object_bad_orientation", TK_BAD_ROLL, p=(math.pi / 3,)),
        ],
        obs_terms=_base_obs() + [ObsTerm("object_state", 13, 1.0, None)],
        gait=GaitParams(with_object=True),
        object_state=ObjectStateObs(),
    )


def student_spec() -> TaskSpec:
    """RandCylinderTransportStudent_SingleBinaryTac: teacher terms with a 10 s episode
    (reference object_transport_student_env_cfg.py:156-201, episode_length_s = 10.0 at :176)."""
    spec = teacher_spec()
    spec.name = "Isaac-RandCylinderTransportStudent_SingleBinaryTac_CNNRNN_Mon-LocoTouch-v1"
    spec.max_episode_length = 500
    return spec


SPECS = {"locomotion": locomotion_spec, "teacher": teacher_spec, "student": student_spec}
