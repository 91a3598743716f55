"""ctypes binding of ``liblocotouch_b200.so`` (the C ABI declared in include/locotouch_b200.h).

The library is the product: there is NO fallback.  If the shared object is missing or a call returns a non-zero
status this module raises -- loudly -- instead of routing around the CUDA path.
PyTorch is used only as the owner of device memory and streams: tensors cross the ABI as raw ``data_ptr()`` values.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "liblocotouch_b200.so")
if os.environ.get("LT_LIB_VARIANT"):  # tuning builds of the same library (python -m locotouch_b200.csrc.build with LT_LIB_VARIANT set)
    LIB_PATH = LIB_PATH[:-3] + f".{os.environ['LT_LIB_VARIANT']}.so"

LT_GATHER_MAX = 12
LT_MAX_REWARD_TERMS = 32
LT_MAX_TERMINATION_TERMS = 8
LT_MAX_OBS_TERMS = 8
LT_MAX_CONTACT_IDS = 8
LT_ERR_UNSUPPORTED = 4  # enum LtStatus
LT_PHASE_REWARDS = 1
LT_PHASE_OBS = 2

f32p = C.c_void_p  # raw device pointers are passed as void*


class LtGatherArgs(C.Structure):
    _fields_ = [
        ("num_tensors", C.c_int),
        ("src", C.c_void_p * LT_GATHER_MAX),
        ("dst", C.c_void_p * LT_GATHER_MAX),
        ("row_len", C.c_int * LT_GATHER_MAX),
    ]


class LtPpoLossArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("A", C.c_int),
        ("mu", C.c_void_p), ("sigma", C.c_void_p), ("value", C.c_void_p), ("actions", C.c_void_p),
        ("old_logp", C.c_void_p), ("old_mu", C.c_void_p), ("old_sigma", C.c_void_p), ("advantages", C.c_void_p),
        ("returns", C.c_void_p), ("old_values", C.c_void_p),
        ("clip_param", C.c_float), ("value_loss_coef", C.c_float), ("entropy_coef", C.c_float),
        ("use_clipped_value_loss", C.c_int),
        ("desired_kl", C.c_float), ("grad_scale", C.c_float),
        ("grad_mu", C.c_void_p), ("grad_value", C.c_void_p), ("grad_sigma", C.c_void_p), ("out", C.c_void_p),
        ("lr_inout", C.c_void_p), ("loss_accum", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class LtPpoHeadsArgs(C.Structure):
    _fields_ = [
        ("loss", LtPpoLossArgs), ("H", C.c_int),
        ("h_actor", C.c_void_p), ("h_critic", C.c_void_p), ("w_actor", C.c_void_p), ("b_actor", C.c_void_p),
        ("w_critic", C.c_void_p), ("b_critic", C.c_void_p), ("g_h_actor", C.c_void_p), ("g_h_critic", C.c_void_p),
    ]


class LtStudentCnnArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("in_channels", C.c_int), ("height", C.c_int), ("width", C.c_int),
        ("channels", C.c_int * 3), ("kernel_sizes", C.c_int * 3), ("pool", C.c_int * 3), ("embedding_dim", C.c_int),
        ("image", C.c_void_p), ("packed", C.c_void_p), ("packed_words", C.c_int),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("wh", C.c_void_p), ("bh", C.c_void_p), ("out", C.c_void_p),
    ]


class LtMlp3Net(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("k0", C.c_int),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("h1", C.c_void_p), ("h2", C.c_void_p), ("h3", C.c_void_p),
    ]


class LtTaxelArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("T", C.c_int),
        ("body_quat_w", C.c_void_p), ("quat_num_bodies", C.c_int), ("quat_body_offset", C.c_int),
        ("net_forces_w", C.c_void_p), ("thresholds", C.c_void_p), ("u_drop", C.c_void_p), ("u_add", C.c_void_p),
        ("p_drop", C.c_float), ("p_add", C.c_float),
        ("seed", C.c_uint64), ("offset", C.c_uint64), ("offset_base", C.c_void_p),
        ("signal", C.c_void_p), ("packed", C.c_void_p), ("normal_forces", C.c_void_p), ("original_contact", C.c_void_p),
        ("delay_ring", C.c_void_p), ("delay_first", C.c_void_p), ("delay_steps", C.c_void_p), ("max_delay", C.c_int),
        ("delayed_signal", C.c_void_p), ("delay_reset", C.c_void_p),
    ]


class LtRewardTerm(C.Structure):
    _fields_ = [("kind", C.c_int), ("weight", C.c_float), ("p", C.c_float * 6)]


class LtTerminationTerm(C.Structure):
    _fields_ = [("kind", C.c_int), ("time_out", C.c_int), ("p", C.c_float * 2), ("num_ids", C.c_int),
                ("body_ids", C.c_int * LT_MAX_CONTACT_IDS)]


class LtObsTerm(C.Structure):
    _fields_ = [("kind", C.c_int), ("dim", C.c_int), ("scale", C.c_float), ("noisy", C.c_int),
                ("n_min", C.c_float), ("n_max", C.c_float)]


class LtGaitParams(C.Structure):
    _fields_ = [
        ("judge_time_threshold", C.c_float), ("air_time_gait_bound", C.c_float), ("contact_time_gait_bound", C.c_float),
        ("async_time_tolerance", C.c_float), ("stance_rwd_scale", C.c_float), ("tolerance_proportion", C.c_float),
        ("rwd_upper_bound", C.c_float), ("rwd_lower_bound", C.c_float), ("vel_tracking_exp_sigma", C.c_float),
        ("task_performance_ratio", C.c_float), ("linear_scale", C.c_float), ("two_step_dt", C.c_float),
        ("async_judge_time_threshold", C.c_float),
        ("encourage_symmetricity", C.c_int), ("with_object", C.c_int),
        ("obj_x_max", C.c_float), ("obj_y_max", C.c_float),
        ("feet_ids", C.c_int * 4),
    ]


class LtGaitState(C.Structure):
    _fields_ = [
        ("last_step_current_air_time", C.c_void_p), ("last_step_current_contact_time", C.c_void_p),
        ("swinging_in_zero_cmd", C.c_void_p), ("valid_last_air_time", C.c_void_p), ("valid_previous_contact", C.c_void_p),
        ("last_velocity_cmd", C.c_void_p), ("step_from_changing_cmd", C.c_void_p),
    ]


class LtTaxelForceArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("T", C.c_int), ("body_quat_w", C.c_void_p), ("quat_num_bodies", C.c_int), ("quat_body_offset", C.c_int),
        ("net_forces_w", C.c_void_p), ("thresholds", C.c_void_p), ("u", C.c_void_p * 7),
        ("seed", C.c_uint64), ("offset", C.c_uint64), ("offset_base", C.c_void_p),
        ("p_drop", C.c_float), ("p_add", C.c_float), ("add_force_noise", C.c_int), ("force_noise_min", C.c_float), ("force_noise_range", C.c_float),
        ("maximal_force", C.c_float), ("level_bin", C.c_float), ("add_level_noise", C.c_int), ("level_noise_min", C.c_float),
        ("level_noise_range", C.c_float), ("out_stride", C.c_int),
        ("contact", C.c_void_p), ("normal_forces", C.c_void_p), ("normalized", C.c_void_p), ("minmax", C.c_void_p), ("discretized", C.c_void_p),
    ]


class LtMdpArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("phases", C.c_int), ("step_dt", C.c_float), ("max_episode_length", C.c_int64),
        ("command", C.c_void_p), ("root_pos_w", C.c_void_p), ("root_quat_w", C.c_void_p), ("root_lin_vel_w", C.c_void_p),
        ("root_ang_vel_w", C.c_void_p), ("root_lin_vel_b", C.c_void_p), ("root_ang_vel_b", C.c_void_p),
        ("projected_gravity_b", C.c_void_p), ("joint_pos", C.c_void_p), ("joint_vel", C.c_void_p), ("joint_acc", C.c_void_p),
        ("applied_torque", C.c_void_p), ("default_joint_pos", C.c_void_p), ("default_joint_vel", C.c_void_p),
        ("soft_joint_pos_limits", C.c_void_p), ("raw_actions", C.c_void_p), ("prev_raw_actions", C.c_void_p),
        ("J", C.c_int),
        ("body_pos_w", C.c_void_p), ("body_lin_vel_w", C.c_void_p), ("num_bodies", C.c_int),
        ("feet_body_ids", C.c_int * 4),
        ("net_forces_w_history", C.c_void_p), ("force_history", C.c_int), ("num_sensor_bodies", C.c_int),
        ("feet_sensor_ids", C.c_int * 4), ("thigh_calf_sensor_ids", C.c_int * 8), ("num_thigh_calf", C.c_int),
        ("current_air_time", C.c_void_p), ("current_contact_time", C.c_void_p), ("last_air_time", C.c_void_p),
        ("episode_length_buf", C.c_void_p),
        ("obj_root_pos_w", C.c_void_p), ("obj_root_quat_w", C.c_void_p), ("obj_root_lin_vel_w", C.c_void_p),
        ("obj_root_ang_vel_w", C.c_void_p), ("obj_projected_gravity_b", C.c_void_p), ("obj_last_contact_time", C.c_void_p),
        ("obj_current_contact_time", C.c_void_p), ("obj_current_air_time", C.c_void_p),
        ("num_reward_terms", C.c_int), ("reward_terms", LtRewardTerm * LT_MAX_REWARD_TERMS),
        ("num_termination_terms", C.c_int), ("termination_terms", LtTerminationTerm * LT_MAX_TERMINATION_TERMS),
        ("gait", LtGaitParams), ("gait_state", LtGaitState),
        ("any_nonzero_cmd_override", C.c_int), ("auto_reset", C.c_int),
        ("reward", C.c_void_p), ("step_reward", C.c_void_p), ("episode_sums", C.c_void_p), ("term_raw", C.c_void_p),
        ("term_masks", C.c_void_p), ("terminated", C.c_void_p), ("time_outs", C.c_void_p), ("dones", C.c_void_p),
        ("episode_log_sums", C.c_void_p),
        ("num_obs_terms", C.c_int), ("obs_terms", LtObsTerm * LT_MAX_OBS_TERMS), ("history_length", C.c_int),
        ("obs_fill", C.c_void_p), ("policy_obs_in", C.c_void_p), ("policy_obs_out", C.c_void_p),
        ("critic_obs_in", C.c_void_p), ("critic_obs_out", C.c_void_p), ("u_obs", C.c_void_p), ("u_obj_euler", C.c_void_p),
        ("seed", C.c_uint64), ("offset", C.c_uint64), ("offset_base", C.c_void_p),
        ("os_n_min", C.c_float * 13), ("os_n_max", C.c_float * 13), ("os_euler_min", C.c_float * 3), ("os_euler_max", C.c_float * 3),
        ("os_scale", C.c_float * 13),
        ("os_non_contact", C.c_float * 13), ("os_last_contact_thr", C.c_float), ("os_current_contact_thr", C.c_float),
        ("any_flag_ws", C.c_void_p), ("tables", C.c_void_p),
        ("act_new", C.c_void_p), ("act_prev_prev_raw", C.c_void_p), ("act_processed", C.c_void_p), ("act_offset", C.c_void_p),
        ("act_clip", C.c_float), ("act_raw_scale", C.c_float), ("act_scale", C.c_float), ("act_reset_on_done", C.c_int),
        ("store_rewards", C.c_void_p), ("store_dones", C.c_void_p), ("store_values", C.c_void_p), ("store_gamma", C.c_float),
    ]


class LtCommandRanges(C.Structure):
    _fields_ = [
        ("ranges", (C.c_double * 2) * 3), ("previous", (C.c_double * 2) * 3), ("equal", C.c_int32 * 3), ("initial_zero_command_steps", C.c_int32),
        ("rel_standing_envs", C.c_double), ("final_initial_zero_command_steps", C.c_int32), ("reserved", C.c_int32),
        ("final_rel_standing_envs", C.c_double), ("lin_forward_bins", C.c_int32), ("ang_forward_bins", C.c_int32),
        ("success_repeat_times_lin", C.c_int32), ("success_repeat_times_ang", C.c_int32),
    ]


LT_CMD_RESET, LT_CMD_COMPUTE = 1, 2
LT_CMD_METRICS = ("error_vel_xy", "error_vel_yaw", "foot_air_time_variance", "foot_step_frequency", "pair_1_step_frequency",
                  "pair_2_step_frequency", "step_air_time", "pair_1_air_time", "pair_2_air_time", "lin_vel_x", "lin_vel_y", "ang_vel_z",
                  "initial_zero_command_steps", "rel_standing_envs")  # enum LtCommandMetric


class LtCommandArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("phases", C.c_int32), ("dt", C.c_float), ("resampling_time_lo", C.c_float), ("resampling_time_hi", C.c_float),
        ("bin_c0", C.c_float), ("bin_c1", C.c_float), ("binary_maximal_command", C.c_int32),
        ("ranges", C.c_void_p), ("vel_command_b", C.c_void_p), ("vel_command_b_buffer", C.c_void_p), ("time_left", C.c_void_p),
        ("command_counter", C.c_void_p), ("is_standing_env", C.c_void_p), ("metrics", C.c_void_p), ("metric_scalars", C.c_void_p),
        ("reset_mask", C.c_void_p), ("reset_extras", C.c_void_p), ("episode_length_buf", C.c_void_p), ("root_lin_vel_b", C.c_void_p),
        ("root_ang_vel_b", C.c_void_p), ("last_air_time", C.c_void_p), ("num_sensor_bodies", C.c_int32), ("feet_ids", C.c_int32 * 4),
        ("gait_valid_last_air_time", C.c_void_p), ("u", C.c_void_p), ("seed", C.c_uint64), ("offset", C.c_uint64), ("offset_base", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class LtVelCurriculumArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("repeat_times_lin", C.c_int32), ("repeat_times_ang", C.c_int32), ("max_distance_bins", C.c_int32),
        ("ranges", C.c_void_p), ("reset_mask", C.c_void_p), ("episode_length_buf", C.c_void_p), ("episode_sums_lin", C.c_void_p),
        ("episode_sums_ang", C.c_void_p), ("env_reseted_lin", C.c_void_p), ("episode_length_buf_lin", C.c_void_p),
        ("episode_reward_sum_lin", C.c_void_p), ("env_reseted_ang", C.c_void_p), ("episode_length_buf_ang", C.c_void_p),
        ("episode_reward_sum_ang", C.c_void_p), ("command_maximum_ranges", C.c_double * 3), ("expansion", C.c_double * 3),
        ("reset_envs_episode_length", C.c_double), ("reward_threshold_lin", C.c_double), ("reward_threshold_ang", C.c_double),
    ]


# name -> (restype, argtypes); must list every symbol include/locotouch_b200.h declares
_SIGNATURES = {
    "lt_abi_version": (C.c_int, []),
    "lt_error_string": (C.c_char_p, [C.c_int]),
    "lt_last_cuda_error": (C.c_char_p, []),
    "lt_struct_size": (C.c_int64, [C.c_int]),
    "lt_gae_workspace_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "lt_gae": (C.c_int, [f32p, f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "lt_gae_scan": (C.c_int, [f32p, f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "lt_adv_normalize": (C.c_int, [f32p, C.c_int64, C.c_void_p, C.c_void_p]),
    "lt_act_sample": (C.c_int, [f32p, f32p, f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "lt_process_actions": (C.c_int, [f32p, C.c_float, C.c_float, C.c_float, f32p, f32p, f32p, f32p, f32p, f32p, f32p, C.c_int64, C.c_void_p]),
    "lt_counter_add": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "lt_store_step": (C.c_int, [f32p, C.c_void_p, C.c_void_p, C.c_void_p, f32p, C.c_float, f32p, C.c_void_p, f32p, f32p, C.c_int, f32p, f32p, C.c_int, C.c_int, C.c_void_p]),
    "lt_gather_rows": (C.c_int, [C.POINTER(LtGatherArgs), C.c_void_p, C.c_int64, C.c_void_p]),
    "lt_ppo_loss_workspace_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "lt_ppo_loss": (C.c_int, [C.POINTER(LtPpoLossArgs), C.c_void_p]),
    "lt_adaptive_lr": (C.c_int, [f32p, C.c_float, C.c_float, f32p, C.c_void_p]),
    "lt_clip_adam_workspace_bytes": (C.c_int64, [C.c_int64]),
    "lt_clip_adam": (C.c_int, [f32p, f32p, f32p, f32p, C.c_int64, f32p, f32p, C.c_float, C.c_double, C.c_double, C.c_float, C.c_float, C.c_float, f32p, C.c_void_p, C.c_int64, C.c_void_p]),
    "lt_peer_sum_clip_adam": (C.c_int, [f32p, C.POINTER(C.c_void_p), C.c_int, f32p, C.c_int, f32p, f32p, C.c_int64, f32p, f32p, C.c_float, C.c_double,
                                        C.c_double, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, f32p, C.c_void_p, C.c_int64, C.c_void_p]),
    "lt_peer_gather_clip_adam": (C.c_int, [f32p, C.POINTER(C.c_void_p), C.c_int, f32p, C.c_int, f32p, f32p, C.c_int64, f32p, f32p, C.c_float, C.c_double,
                                        C.c_double, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, f32p, C.c_void_p, C.c_int64, C.c_void_p]),
    "lt_peer_reduce_scatter": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    "lt_bias_act_bwd_workspace_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "lt_bias_act_bwd": (C.c_int, [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]),
    "lt_taxel_synth": (C.c_int, [C.POINTER(LtTaxelArgs), C.c_void_p]),
    "lt_tactile_delay": (C.c_int, [f32p, C.c_void_p, C.c_void_p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "lt_mdp_step": (C.c_int, [C.POINTER(LtMdpArgs), C.c_void_p]),
    "lt_trajectory_index": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "lt_split_pad_trajectories": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p]),
    "lt_unpad_trajectories": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p]),
    "lt_dagger_step": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_int64, C.c_int] + [C.c_void_p] * 7),
    "lt_pack_trajectories": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "lt_taxel_forces": (C.c_int, [C.POINTER(LtTaxelForceArgs), C.c_void_p]),
    "lt_linear_bias_act_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "lt_linear_bias_act": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p, C.c_int64, C.c_void_p]),
    "lt_dgrad_act_bwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p, C.c_int64, C.c_void_p]),
    "lt_ppo_heads_workspace_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "lt_ppo_heads_loss": (C.c_int, [C.POINTER(LtPpoHeadsArgs), C.c_void_p]),
    "lt_student_cnn_forward": (C.c_int, [C.POINTER(LtStudentCnnArgs), C.c_void_p]),
    "lt_contact_sensor_update": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p] * 5 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "lt_act_heads": (C.c_int, [C.c_void_p] * 13 + [C.c_int] * 3 + [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "lt_wgrad_splitk": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p]),
    "lt_mlp3_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "lt_wgrad_splitk_pair": (C.c_int, [C.c_void_p] * 8 + [C.c_int] * 3 + [C.c_void_p]),
    "lt_mdp_tables_len": (C.c_int, [C.POINTER(LtMdpArgs)]),
    "lt_mdp_build_tables": (C.c_int, [C.POINTER(LtMdpArgs), C.POINTER(C.c_int32), C.c_int]),
    "lt_mdp_reset": (C.c_int, [C.POINTER(LtGaitState), f32p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "lt_pad_trajectories": (C.c_int, [f32p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, f32p, C.c_void_p, C.c_void_p]),
    "lt_gemm_backend": (C.c_char_p, []),
    "lt_masked_mse_workspace_bytes": (C.c_int64, [C.c_int64]),
    "lt_command_workspace_bytes": (C.c_int64, [C.c_int]),
    "lt_command_step": (C.c_int, [C.POINTER(LtCommandArgs), C.c_void_p]),
    "lt_vel_curriculum": (C.c_int, [C.POINTER(LtVelCurriculumArgs), C.c_void_p]),
    "lt_masked_mse": (C.c_int, [f32p, f32p, C.c_void_p, C.c_int64, C.c_int, f32p, f32p, C.c_void_p, C.c_int64, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class LocoTouchLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads the shared object (once).  Raises if it has not been built: there is no CPU / eager fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LocoTouchLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m locotouch_b200.csrc.build` "
            "(or __graft_entry__.build()).  locotouch_b200 has no fallback path."
        )
    handle = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as e:  # pragma: no cover
            raise LocoTouchLibraryError(f"{LIB_PATH} does not export {name}; rebuild the library") from e
        fn.restype = restype
        fn.argtypes = argtypes
    if handle.lt_abi_version() != 1:
        raise LocoTouchLibraryError("ABI version mismatch between _C.py and liblocotouch_b200.so")
    for which, struct in enumerate((LtGatherArgs, LtPpoLossArgs, LtTaxelArgs, LtMdpArgs, LtGaitState, LtGaitParams, LtTaxelForceArgs,
                                    LtCommandRanges, LtCommandArgs, LtVelCurriculumArgs, LtPpoHeadsArgs, LtStudentCnnArgs, LtMlp3Net)):
        if handle.lt_struct_size(which) != C.sizeof(struct):
            raise LocoTouchLibraryError(
                f"struct layout mismatch for {struct.__name__}: C {handle.lt_struct_size(which)} vs ctypes {C.sizeof(struct)}")
    _lib = handle
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        handle = lib()
        msg = handle.lt_error_string(status).decode()
        cuda = handle.lt_last_cuda_error().decode()
        raise LocoTouchLibraryError(f"{what or 'locotouch_b200 call'} failed: {msg}" + (f" [{cuda}]" if cuda and status == 2 else ""))


def ptr(t: torch.Tensor | None, dtype: torch.dtype | None = None, name: str = "tensor") -> int | None:
    """Raw device pointer of a contiguous CUDA tensor (``None`` passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise LocoTouchLibraryError(f"{name} must live on a CUDA device (got {t.device}); locotouch_b200 has no CPU path")
    if not t.is_contiguous():
        raise LocoTouchLibraryError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise LocoTouchLibraryError(f"{name} must be {dtype} (got {t.dtype})")
    if t.device.index != torch.cuda.current_device():
        # every launch goes to the CURRENT device's current stream (and sizes its grid from that device): a tensor of another GPU
        # would be dereferenced on the wrong device
        raise LocoTouchLibraryError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                                    "call torch.cuda.set_device(...) / use `with torch.cuda.device(...)` around locotouch_b200 calls")
    return t.data_ptr()


def gemm_backend() -> str:
    """Which implementation of the fused GEMMs (K12 ...) this build of the library carries; "stub" = built without them."""
    return lib().lt_gemm_backend().decode()


def current_stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# launch counter: bench.py reports how many of OUR kernels-launching ABI calls ran inside the timed region
launch_count = 0


def count_launches(n: int = 1):
    global launch_count
    launch_count += n
