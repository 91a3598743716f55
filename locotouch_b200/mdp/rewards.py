"""Reward terms with the reference's names and manager-term signatures (reference locotouch/mdp/rewards.py).

``f(env, **params) -> Tensor[num_envs]`` / class terms with ``__init__(cfg, env)``, ``reset(env_ids)``, ``__call__``.
Every callable returns its row of ONE fused launch per env step (``lt_mdp_step``, K1); nothing is computed in Python.
Values are float32 (the reference returns int64 / bool for the counting terms; RewardManager multiplies by a float weight
either way).  Rows are views into the fused result buffer: treat them as read-only.
"""
from __future__ import annotations

import torch

from . import task_spec as TS
from ._fusion import cache_for, reset_terms, reward_term

__all__ = [
    "track_lin_vel_xy_pst", "track_ang_vel_z_pst", "foot_slipping_ngt", "foot_dragging_ngt", "AdaptiveSymmetricGaitReward",
    "AdaptiveSymmetricGaitRewardwithObject", "track_base_height_ngt", "base_z_velocity_ngt", "base_roll_pitch_velocity_ngt",
    "base_roll_pitch_angle_ngt", "joint_position_limit_ngt", "joint_position_ngt", "joint_velocity_ngt", "joint_acceleration_ngt",
    "joint_torque_ngt", "action_rate_ngt", "thigh_calf_collision_ngt", "object_relative_xy_position_ngt",
    "object_relative_xy_velocity_ngt", "object_relative_z_velocity_ngt", "object_relative_roll_pitch_angle_ngt",
    "object_relative_roll_pitch_velocity_ngt", "object_relative_roll_angle_ngt", "object_relative_roll_velocity_ngt",
    "object_relative_yaw_angle_ngt", "object_dangerous_state_ngt", "object_lose_contact_ngt",
]


# ----------------- Velocity Tracking Task (reference rewards.py:15-27)
def track_lin_vel_xy_pst(env, sigma: float = 0.25, command_name: str = "base_velocity", asset_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_TRACK_LIN_VEL_XY, (sigma,))


def track_ang_vel_z_pst(env, sigma: float = 0.25, command_name: str = "base_velocity", asset_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_TRACK_ANG_VEL_Z, (sigma,))


# ----------------- Foot Slipping and Dragging (reference rewards.py:31-56)
def foot_slipping_ngt(env, threshold: float = 1.0, asset_cfg=None, sensor_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_FOOT_SLIP, (threshold,))


def foot_dragging_ngt(env, asset_cfg=None, height_threshold: float = 0.025, foot_vel_xy_threshold: float = 0.1) -> torch.Tensor:
    return reward_term(env, TS.RK_FOOT_DRAG, (height_threshold, foot_vel_xy_threshold))


# ----------------- Gait (reference rewards.py:60-392)
class AdaptiveSymmetricGaitReward:
    """Class term; the seven state arrays of reference rewards.py:96-105 live in the fused object and are exposed under the
    reference's attribute names (``valid_last_air_time`` is what ``commands.py:399-417`` reads for gait logging)."""

    _with_object = False

    def __init__(self, cfg, env):
        self.cfg = cfg
        self._env = env
        names = cfg.params["synced_feet_pair_names"]
        if len(names) != 2 or len(names[0]) != 2 or len(names[1]) != 2:
            raise ValueError("This reward only supports gaits with two pairs of synchronized feet, like trotting.")
        self._fused = cache_for(env).fused
        gp = self._fused.spec.gait
        for key, val in cfg.params.items():
            if hasattr(gp, key) and key != "synced_feet_pair_names" and abs(float(getattr(gp, key)) - float(val)) > 1e-12:
                raise ValueError(f"gait parameter {key}={val} differs from the fused table ({getattr(gp, key)}); set env.lt_task_spec first")
        if bool(gp.with_object) != self._with_object:
            raise ValueError("gait term class and scene disagree about the transported object")

    num_envs = property(lambda self: self._env.num_envs)
    device = property(lambda self: self._env.device)
    last_step_current_air_time = property(lambda self: self._fused.last_step_current_air_time)
    last_step_current_contact_time = property(lambda self: self._fused.last_step_current_contact_time)
    swinging_in_zero_cmd = property(lambda self: self._fused.swinging_in_zero_cmd)
    valid_last_air_time = property(lambda self: self._fused.valid_last_air_time)
    valid_previous_contact = property(lambda self: self._fused.valid_previous_contact)
    last_velocity_cmd = property(lambda self: self._fused.last_velocity_cmd)
    step_from_changing_cmd = property(lambda self: self._fused.step_from_changing_cmd)

    def reset(self, env_ids=None):
        reset_terms(self._env, env_ids)

    def __call__(self, env, **params) -> torch.Tensor:
        return reward_term(env, TS.RK_GAIT)


class AdaptiveSymmetricGaitRewardwithObject(AdaptiveSymmetricGaitReward):
    _with_object = True


# ----------------- Regularization (reference rewards.py:398-466)
def track_base_height_ngt(env, target_height: float = 0.42, asset_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_BASE_HEIGHT, (target_height,))


def base_z_velocity_ngt(env, asset_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_BASE_Z_VEL)


def base_roll_pitch_velocity_ngt(env, asset_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_BASE_RP_VEL)


def base_roll_pitch_angle_ngt(env, asset_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_BASE_RP_ANGLE)


def joint_position_limit_ngt(env, asset_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_JOINT_POS_LIMIT)


def joint_position_ngt(env, asset_cfg=None, stand_still_scale: float = 5.0, velocity_threshold: float = 0.3) -> torch.Tensor:
    return reward_term(env, TS.RK_JOINT_POS, (stand_still_scale, velocity_threshold))


def joint_velocity_ngt(env, asset_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_JOINT_VEL)


def joint_acceleration_ngt(env, asset_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_JOINT_ACC)


def joint_torque_ngt(env, asset_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_JOINT_TORQUE)


def action_rate_ngt(env) -> torch.Tensor:
    return reward_term(env, TS.RK_ACTION_RATE)


def thigh_calf_collision_ngt(env, threshold: float = 0.1, sensor_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_THIGH_CALF_COLLISION, (threshold,))


# ----------------- Object Transport (reference rewards.py:469-604)
def object_relative_xy_position_ngt(env, robot_cfg=None, object_cfg=None, work_only_when_cmd: int = 0) -> torch.Tensor:
    return reward_term(env, TS.RK_OBJ_XY_POS, (float(bool(work_only_when_cmd)),))


def object_relative_xy_velocity_ngt(env, robot_cfg=None, object_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_OBJ_XY_VEL)


def object_relative_z_velocity_ngt(env, robot_cfg=None, object_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_OBJ_Z_VEL)


def object_relative_roll_pitch_angle_ngt(env, robot_cfg=None, object_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_OBJ_RP_ANGLE)


def object_relative_roll_pitch_velocity_ngt(env, robot_cfg=None, object_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_OBJ_RP_VEL)


def object_relative_roll_angle_ngt(env, robot_cfg=None, object_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_OBJ_ROLL_ANGLE)


def object_relative_roll_velocity_ngt(env, robot_cfg=None, object_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_OBJ_ROLL_VEL)


def object_relative_yaw_angle_ngt(env, robot_cfg=None, object_cfg=None, work_only_when_cmd: int = 0) -> torch.Tensor:
    return reward_term(env, TS.RK_OBJ_YAW, (float(bool(work_only_when_cmd)),))


def object_dangerous_state_ngt(env, robot_cfg=None, object_cfg=None, x_max=None, y_max=None, z_min=None, roll_pitch_max=None,
                               vel_xy_max=None) -> torch.Tensor:
    if x_max is None or y_max is None or z_min is None:
        raise ValueError("the fused object_dangerous_state term needs x_max, y_max and z_min")
    return reward_term(env, TS.RK_OBJ_DANGER, (x_max, y_max, z_min, -1.0 if roll_pitch_max is None else roll_pitch_max,
                                               -1.0 if vel_xy_max is None else vel_xy_max))


def object_lose_contact_ngt(env, object_cfg=None, sensor_cfg=None) -> torch.Tensor:
    return reward_term(env, TS.RK_OBJ_LOSE_CONTACT)
