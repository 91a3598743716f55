/* locotouch_b200 -- C ABI of the B200-native LocoTouch hot path (liblocotouch_b200.so, sm_100a).
 *
 * Every entry point is stateless: raw DEVICE pointers + sizes + scalar parameters + the CUDA stream to launch on
 * (passed as void* == cudaStream_t; 0 is the legacy default stream).  No allocation happens inside the library;
 * workspaces are passed in (query the size with the matching *_workspace_bytes function).  All functions return an
 * int status (LT_OK == 0); lt_error_string() maps it to text.  Launches are asynchronous and capturable into CUDA
 * graphs.  All floating point is IEEE fp32 unless stated; masks are uint8 (0/1); dones uint8; episode_length int64.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference repository root).
 */
#ifndef LOCOTOUCH_B200_H
#define LOCOTOUCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LT_ABI_VERSION 1

enum LtStatus {
  LT_OK = 0,
  LT_ERR_INVALID_ARG = 1,  /* null pointer, non-positive size, unsupported dimension */
  LT_ERR_CUDA = 2,         /* kernel launch failed: see lt_last_cuda_error() */
  LT_ERR_WORKSPACE = 3,    /* workspace too small */
  LT_ERR_UNSUPPORTED = 4
};

int lt_abi_version(void);
const char* lt_error_string(int status);
/* Text of the last CUDA error seen by this thread inside the library ("" if none). */
const char* lt_last_cuda_error(void);
/* sizeof() of the argument structs as compiled, so that a foreign-language binding can verify its own layout:
 * which = 10 LtPpoHeadsArgs, 11 LtStudentCnnArgs, 12 LtMlp3Net, 0 LtGatherArgs, 1 LtPpoLossArgs, 2 LtTaxelArgs, 3 LtMdpArgs, 4 LtGaitState, 5 LtGaitParams, 6 LtTaxelForceArgs,
 * 7 LtCommandRanges, 8 LtCommandArgs, 9 LtVelCurriculumArgs; -1 otherwise. */
int64_t lt_struct_size(int which);

/* ------------------------------------------------------------------------------------------------------------------
 * K4  GAE returns + advantage normalisation
 * replaces  loco_rl/loco_rl/storage/rollout_storage.py:152-174  RolloutStorage.compute_returns
 *   rewards, values [T,N] f32 ; dones [T,N] u8 ; last_values [N] f32  ->  returns, advantages [T,N] f32
 *   delta = r + (1-d)*gamma*V' - V ; A = delta + (1-d)*gamma*lam*A' ; R = A + V ; adv = R - V
 *   normalize != 0:  adv = (adv - mean) / (std_unbiased + 1e-8) over all T*N elements.
 * lt_gae = lt_gae_scan + lt_adv_normalize on the same stream.  Multi-GPU callers run lt_gae_scan, all-reduce the
 * three doubles in `stats` (sum, sum of squares, count) and then call lt_adv_normalize.
 * ------------------------------------------------------------------------------------------------------------------ */
int64_t lt_gae_workspace_bytes(int T, int N);
int lt_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
           float* returns, float* advantages, int T, int N, float gamma, float lam, int normalize,
           void* workspace, int64_t workspace_bytes, void* stream);
/* stats: 4 doubles on the device = {sum(adv), sum(adv^2), count, unused}; overwritten. */
int lt_gae_scan(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                float* returns, float* advantages, int T, int N, float gamma, float lam,
                double* stats, void* workspace, int64_t workspace_bytes, void* stream);
int lt_adv_normalize(float* advantages, int64_t count, const double* stats, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K3  act epilogue + transition store
 * replaces  loco_rl/loco_rl/modules/actor_critic.py:105-123 (Normal sample / log_prob) as used by
 *           loco_rl/loco_rl/algorithms/ppo.py:129-141 (PPO.act), and
 *           ppo.py:143-170 (process_env_step: timeout bootstrap) + rollout_storage.py:80-107 (add_transitions)
 * lt_act_sample:   actions = mu + sigma*eps ; logp = sum_j[-(a-mu)^2/(2 sigma^2) - log sigma - log sqrt(2 pi)]
 *                  and writes mu / sigma(expanded) rows; all outputs may point straight into the rollout slot.
 *                  eps == NULL -> in-kernel Philox4x32-10 normal draws keyed by (seed, offset, sample, j).
 * lt_store_step:   rewards_out = r + gamma*V*time_out ; dones_out = (u8) dones ; optional row copies of obs /
 *                  critic obs (skipped when src == dst or src == NULL).
 * ------------------------------------------------------------------------------------------------------------------ */
int lt_act_sample(const float* mu, const float* sigma /*[A]*/, const float* eps /*[N,A] or NULL*/,
                  float* actions, float* logp, float* mu_out, float* sigma_out, int N, int A,
                  uint64_t seed, uint64_t offset, const int64_t* offset_base /*device counter added to offset, or NULL*/,
                  void* stream);
/* *counter += inc on the stream: advances the device-resident step counter that `offset_base` arguments point at, so
 * that a captured CUDA graph draws fresh random numbers (and hands the right step index on) at every replay. */
int lt_counter_add(int64_t* counter, int64_t inc, void* stream);
int lt_store_step(const float* rewards, const int64_t* dones_i64, const uint8_t* dones_u8, const uint8_t* time_outs,
                  const float* values, float gamma, float* rewards_out, uint8_t* dones_out,
                  const float* obs, float* obs_out, int obs_dim,
                  const float* critic_obs, float* critic_obs_out, int critic_obs_dim, int N, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K0  action pre-processing
 * replaces  locotouch/mdp/actions.py:30-44  JointPositionActionPrevPrev.process_actions (+ [IL] JointPositionAction):
 *   prev_prev_raw = prev_raw ; prev_raw = raw ; prev_prev_processed = prev_processed ; prev_processed = processed
 *   raw = clamp(actions, -clip, clip) * raw_scale   (clip <= 0: no clipping)
 *   processed = raw * scale + offset[N,J]
 * All state tensors are [N,J]; the two "processed" history buffers may be NULL.
 * ------------------------------------------------------------------------------------------------------------------ */
int lt_process_actions(const float* actions, float clip, float raw_scale, float scale, const float* offset,
                       float* raw, float* prev_raw, float* prev_prev_raw,
                       float* processed, float* prev_processed, float* prev_prev_processed, int64_t count, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K5  mini-batch gather
 * replaces  rollout_storage.py:186-243  RolloutStorage.mini_batch_generator  (9 advanced-index gathers)
 * Gathers `count` rows given by indices[count] (int64, into the flattened [T*N] axis) from up to LT_GATHER_MAX
 * row-major sources in ONE launch.  The permutation is drawn once per update() (rollout_storage.py:189) and reused
 * by every epoch, so a caller may gather the whole permuted rollout once and slice it afterwards.
 * ------------------------------------------------------------------------------------------------------------------ */
#define LT_GATHER_MAX 12
typedef struct {
  int num_tensors;
  const float* src[LT_GATHER_MAX];
  float* dst[LT_GATHER_MAX];
  int row_len[LT_GATHER_MAX]; /* floats per row */
} LtGatherArgs;
int lt_gather_rows(const LtGatherArgs* args, const int64_t* indices, int64_t count, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K6  fused PPO loss (forward + analytic backward)
 * replaces  loco_rl/loco_rl/algorithms/ppo.py:252-302 (log-prob, entropy, KL, adaptive LR, clipped surrogate,
 *           clipped value loss, total loss) and the autograd backward of those lines up to mu / sigma / value.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int B;                   /* mini-batch rows */
  int A;                   /* action dim, multiple of 4, <= 64 */
  const float* mu;         /* [B,A] actor output */
  const float* sigma;      /* [A]   state-independent std parameter */
  const float* value;      /* [B]   critic output */
  const float* actions;    /* [B,A] */
  const float* old_logp;   /* [B] */
  const float* old_mu;     /* [B,A] */
  const float* old_sigma;  /* [B,A] */
  const float* advantages; /* [B] */
  const float* returns;    /* [B] */
  const float* old_values; /* [B] */
  float clip_param, value_loss_coef, entropy_coef;
  int use_clipped_value_loss;
  float desired_kl;        /* adaptive schedule when > 0 and lr_inout != NULL (ppo.py:264-281) */
  float grad_scale;        /* multiplies every gradient (1 for one process; 1 also for DDP: the all-reduce averages) */
  float* grad_mu;          /* [B,A]  dLoss/dmu */
  float* grad_value;       /* [B]    dLoss/dvalue */
  float* grad_sigma;       /* [A]    dLoss/dsigma (overwritten, deterministic two-stage reduction) */
  float* out;              /* [8] = {loss, surrogate, value_loss, entropy_mean, kl_mean, lr_after, 0, 0} */
  float* lr_inout;         /* device scalar learning rate or NULL */
  float* loss_accum;       /* optional [4] running sums {value_loss, surrogate, entropy, count} (ppo.py:361-363) or NULL */
  void* workspace;
  int64_t workspace_bytes;
} LtPpoLossArgs;
int64_t lt_ppo_loss_workspace_bytes(int B, int A);
int lt_ppo_loss(const LtPpoLossArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K16  output heads of both MLPs + PPO loss + head dgrad in ONE pass over the mini-batch
 * replaces  the two head nn.Linear layers of loco_rl/loco_rl/modules/actor_critic.py:33-56 as PPO.update evaluates them
 *           (algorithms/ppo.py:252-255: `act` / `evaluate` on the mini-batch), the loss block ppo.py:256-302 (= K6) and autograd's
 *           backward of the heads down to the pre-activation of the last hidden layer (ppo.py:350): mu = h_a W_a^T + b_a,
 *           V = h_c W_c^T + b_c, the loss and its derivatives, g_h = (dL/dout . W) * elu'(h) for both networks.
 * `loss` is the K6 argument block with mu / value turned into optional OUTPUTS (may be NULL) and grad_mu [B,A] / grad_value [B]
 * required outputs (the head weight / bias gradients are K15's, from these two tensors and h).  h_* are the post-ELU activations
 * [B,H] of the last hidden layers (both of width H, H % 128 == 0, H <= 256; A % 4 == 0, A <= 16; LT_ERR_UNSUPPORTED otherwise),
 * w_actor [A,H], b_actor [A], w_critic [1,H], b_critic [1]; g_h_* [B,H] are overwritten.  fp32 FMA arithmetic throughout.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  LtPpoLossArgs loss;
  int H;
  const float* h_actor;
  const float* h_critic;
  const float* w_actor;
  const float* b_actor;
  const float* w_critic;
  const float* b_critic;
  float* g_h_actor;
  float* g_h_critic;
} LtPpoHeadsArgs;
int64_t lt_ppo_heads_workspace_bytes(int B, int A);
int lt_ppo_heads_loss(const LtPpoHeadsArgs* args, void* stream);
/* lr = max(1e-5, lr/1.5) if kl > 2*desired ; lr = min(1e-2, lr*1.5) if 0 < kl < desired/2.  kl = *kl_sum * kl_scale. */
int lt_adaptive_lr(const float* kl_sum, float kl_scale, float desired_kl, float* lr_inout, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K7  global-norm clip + Adam over one flat parameter buffer
 * replaces  ppo.py:350-353  nn.utils.clip_grad_norm_(params, max_norm) ; optim.Adam.step()  (and AdamW for the
 *           student, locotouch/distill/student.py:82,151, with weight_decay > 0)
 *   g *= grad_scale ; total = ||g||_2 ; g *= min(1, max_norm/(total+1e-6)) (skipped when max_norm <= 0)
 *   p *= 1 - lr*weight_decay ; m = m + (g-m)(1-b1) ; v = b2 v + (1-b2) g^2
 *   p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * lr and the step counter live on the device so that a captured graph can be replayed.
 * ------------------------------------------------------------------------------------------------------------------ */
int64_t lt_clip_adam_workspace_bytes(int64_t n);
int lt_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                 const float* lr, float* step_inout, float max_grad_norm, double beta1, double beta2, float eps,
                 float weight_decay, float grad_scale, float* grad_norm_out /*or NULL*/,
                 void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K9  fused ELU backward + bias gradient (one pass over the [B, n] layer gradient)
 * replaces  the autograd kernels behind every Linear+ELU layer of loco_rl/loco_rl/modules/actor_critic.py:33-56 and
 *           loco_rl/loco_rl/models/mlp.py:4-25: elu_backward and the bias-gradient column sum.
 *   grad_pre = grad_out * (act_out > 0 ? 1 : act_out + alpha)   (act_out == NULL: no activation, grad_pre = grad_out)
 *   bias_grad[j] = sum_b grad_pre[b, j]                          (deterministic two-stage reduction)
 * grad_pre may alias grad_out (in place) or be NULL when only the column sums are wanted.
 * ------------------------------------------------------------------------------------------------------------------ */
int64_t lt_bias_act_bwd_workspace_bytes(int B, int n);
int lt_bias_act_bwd(const float* grad_out, const float* act_out, float* grad_pre, float* bias_grad, int B, int n, float alpha,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K2  binary taxel synthesis (+ fused delay line)
 * replaces  locotouch/mdp/observations.py:154-159,166-199,281-308  BinaryTactileSignals.__call__
 *           locotouch/distill/tactile_recorder.py:4-34             TactileRecorder (lt_tactile_delay)
 *   F_n = -(R(q_taxel)^T F_w).z ; C = F_n > thr ; C' = C & !(U_drop < p_drop) ; out = C' | (U_add < p_add)
 *   signal [N, 2*T] f32 = (out, out) ; packed [N, ceil(T/32)] u32 warp-ballot bitmap (bit t%32 of word t/32).
 *   u_drop / u_add == NULL -> in-kernel Philox uniforms keyed by (seed, offset, env, taxel).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int N;                      /* envs */
  int T;                      /* taxels per env (17*13 = 221) */
  const float* body_quat_w;   /* [N, quat_stride_bodies, 4] (w,x,y,z); taxel t is body quat_body_offset + t */
  int quat_num_bodies;        /* bodies per env in body_quat_w */
  int quat_body_offset;
  const float* net_forces_w;  /* [N, T, 3] */
  const float* thresholds;    /* [N, T] */
  const float* u_drop;        /* [N, T] or NULL */
  const float* u_add;         /* [N, T] or NULL */
  float p_drop, p_add;
  uint64_t seed, offset;
  const int64_t* offset_base; /* optional device counter added to offset */
  float* signal;              /* [N, 2T] or NULL */
  uint32_t* packed;           /* [N, ceil(T/32)] or NULL */
  float* normal_forces;       /* [N, T] or NULL  (original_normal_forces side buffer) */
  uint8_t* original_contact;  /* [N, T] or NULL  (original_contact_taxels side buffer) */
  /* fused delay line on the packed bitmap (all NULL/0 to disable): ring [N, max_delay, words] */
  uint32_t* delay_ring;
  uint8_t* delay_first;       /* [N] 1 = first frame after reset -> fill every slot */
  const int64_t* delay_steps; /* [N] slot to read */
  int max_delay;
  float* delayed_signal;      /* [N, 2T] */
  const uint8_t* delay_reset; /* [N] or NULL: envs reset since the previous frame (the dones of the previous env step, e.g. a row of
                                 RolloutStorage.dones): treated like delay_first, read only -- saves the separate flag |= dones pass */
} LtTaxelArgs;
int lt_taxel_synth(const LtTaxelArgs* args, void* stream);
/* Force-valued tactile encodings (reference locotouch/mdp/observations.py:166-237; classes NormalizedTactileSignals,
 * DiscreteTactileSignals, CotinuousTactileSignals, ProcessedTactileSignals :311-429).  Same inputs as lt_taxel_synth.
 * u[0..6]: optional explicit [N, T] uniforms standing for the reference's rand_like draws (drop, drop force, add, add force,
 * force noise, too-small replacement, level noise); u[0] == NULL -> Philox (seed, offset [+ *offset_base]).
 * The derived floats are computed by the caller in double and rounded once, like the Python scalars of the reference:
 * force_noise_range = n_max - n_min, level_bin = 1 / total_levels, level_noise_range = level_n_max - level_n_min.
 * Outputs (each optional): element (n, t) at ptr[n * out_stride + t], so the channels of one [N, C, T] tensor can be passed as
 * ptr + c * T with out_stride = C * T.  contact is written as 0.0 / 1.0. */
typedef struct {
  int N, T;
  const float* body_quat_w; int quat_num_bodies, quat_body_offset;
  const float* net_forces_w;  /* [N, T, 3] */
  const float* thresholds;    /* [N, T] */
  const float* u[7];
  uint64_t seed, offset; const int64_t* offset_base;
  float p_drop, p_add;
  int add_force_noise; float force_noise_min, force_noise_range;
  float maximal_force;
  float level_bin; int add_level_noise; float level_noise_min, level_noise_range;
  int out_stride;
  float* contact; float* normal_forces; float* normalized; float* minmax; float* discretized;
} LtTaxelForceArgs;
int lt_taxel_forces(const LtTaxelForceArgs* args, void* stream);
/* Generic fp32 delay line: ring [N, max_delay, D]; shift in `signal`, first-frame fill, out[n] = ring[n, delay[n]]. */
int lt_tactile_delay(float* ring, uint8_t* first, const int64_t* delay_steps, const float* signal, float* out,
                     int N, int max_delay, int D, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K1  fused MDP step: terminations -> rewards (incl. stateful gait term) -> [auto reset] -> observations
 * replaces  locotouch/mdp/rewards.py:15-604, locotouch/mdp/terminations.py:10-23,
 *           locotouch/mdp/observations.py:38-91 and the IsaacLab manager loops that call them
 *           (RewardManager.compute / TerminationManager.compute / ObservationManager.compute).
 * ------------------------------------------------------------------------------------------------------------------ */
#define LT_MAX_REWARD_TERMS 32
#define LT_MAX_TERMINATION_TERMS 8
#define LT_MAX_OBS_TERMS 8
#define LT_MAX_CONTACT_IDS 8

enum LtRewardKind {
  LT_RK_ALIVE = 0, LT_RK_TRACK_LIN_VEL_XY, LT_RK_TRACK_ANG_VEL_Z, LT_RK_FOOT_SLIP, LT_RK_FOOT_DRAG, LT_RK_GAIT,
  LT_RK_BASE_HEIGHT, LT_RK_BASE_Z_VEL, LT_RK_BASE_RP_ANGLE, LT_RK_BASE_RP_VEL, LT_RK_JOINT_POS_LIMIT,
  LT_RK_JOINT_POS, LT_RK_JOINT_ACC, LT_RK_JOINT_VEL, LT_RK_JOINT_TORQUE, LT_RK_ACTION_RATE,
  LT_RK_THIGH_CALF_COLLISION, LT_RK_OBJ_XY_POS, LT_RK_OBJ_XY_VEL, LT_RK_OBJ_LOSE_CONTACT, LT_RK_OBJ_Z_VEL,
  LT_RK_OBJ_RP_ANGLE, LT_RK_OBJ_RP_VEL, LT_RK_OBJ_ROLL_ANGLE, LT_RK_OBJ_ROLL_VEL, LT_RK_OBJ_YAW, LT_RK_OBJ_DANGER,
  LT_RK_COUNT
};
enum LtTerminationKind {
  LT_TK_TIME_OUT = 0, LT_TK_BAD_ORIENTATION, LT_TK_ROOT_HEIGHT, LT_TK_ILLEGAL_CONTACT, LT_TK_OBJECT_BELOW_ROBOT,
  LT_TK_BAD_ROLL
};
enum LtObsKind {
  LT_OK_COMMAND = 0, LT_OK_BASE_ANG_VEL, LT_OK_PROJECTED_GRAVITY, LT_OK_JOINT_POS_REL, LT_OK_JOINT_VEL_REL,
  LT_OK_LAST_ACTION, LT_OK_OBJECT_STATE
};
enum LtMdpPhase { LT_PHASE_REWARDS = 1, LT_PHASE_OBS = 2 };

typedef struct { int kind; float weight; float p[6]; } LtRewardTerm;
typedef struct { int kind; int time_out; float p[2]; int num_ids; int body_ids[LT_MAX_CONTACT_IDS]; } LtTerminationTerm;
typedef struct { int kind; int dim; float scale; int noisy; float n_min, n_max; } LtObsTerm;

typedef struct {
  /* rewards.py:168-188 parameters of AdaptiveSymmetricGaitReward */
  float judge_time_threshold, air_time_gait_bound, contact_time_gait_bound, async_time_tolerance;
  float stance_rwd_scale, tolerance_proportion, rwd_upper_bound, rwd_lower_bound;
  float vel_tracking_exp_sigma, task_performance_ratio, linear_scale, two_step_dt;
  float async_judge_time_threshold; /* float(judge_time_threshold + async_time_tolerance), rewards.py:67 */
  int encourage_symmetricity; /* bool */
  int with_object;            /* AdaptiveSymmetricGaitRewardwithObject */
  float obj_x_max, obj_y_max;
  int feet_ids[4];            /* contact-sensor body ids in gait order (pair0[0], pair0[1], pair1[0], pair1[1]) */
} LtGaitParams;

typedef struct {
  /* rewards.py:96-105 state arrays, read AND written */
  float* last_step_current_air_time;     /* [N,4] */
  float* last_step_current_contact_time; /* [N,4] */
  uint8_t* swinging_in_zero_cmd;         /* [N,4] */
  float* valid_last_air_time;            /* [N,4] */
  uint8_t* valid_previous_contact;       /* [N,4] */
  float* last_velocity_cmd;              /* [N,3] */
  float* step_from_changing_cmd;         /* [N]   */
} LtGaitState;

typedef struct {
  int N;
  int phases;                    /* LT_PHASE_REWARDS | LT_PHASE_OBS */
  float step_dt;
  int64_t max_episode_length;
  /* ---- inputs: exactly the IsaacLab tensors the reference terms read, contiguous row-major ---- */
  const float* command;          /* [N,3]  command_manager.get_command("base_velocity") */
  const float* root_pos_w;       /* [N,3] */
  const float* root_quat_w;      /* [N,4] wxyz (object tasks only) */
  const float* root_lin_vel_w;   /* [N,3] (object tasks only) */
  const float* root_ang_vel_w;   /* [N,3] (object tasks only) */
  const float* root_lin_vel_b;   /* [N,3] */
  const float* root_ang_vel_b;   /* [N,3] */
  const float* projected_gravity_b; /* [N,3] */
  const float* joint_pos;        /* [N,J] */
  const float* joint_vel;        /* [N,J] */
  const float* joint_acc;        /* [N,J] */
  const float* applied_torque;   /* [N,J] */
  const float* default_joint_pos;/* [N,J] */
  const float* default_joint_vel;/* [N,J] */
  const float* soft_joint_pos_limits; /* [N,J,2] */
  const float* raw_actions;      /* [N,J] action_manager.get_term("joint_pos").raw_actions */
  const float* prev_raw_actions; /* [N,J] */
  int J;                         /* joints (12) */
  const float* body_pos_w;       /* [N,num_bodies,3] */
  const float* body_lin_vel_w;   /* [N,num_bodies,3] */
  int num_bodies;
  int feet_body_ids[4];          /* articulation body ids of ".*foot" (slip / dragging) */
  const float* net_forces_w_history; /* [N,H,num_sensor_bodies,3] */
  int force_history;             /* H (3) */
  int num_sensor_bodies;         /* 17 */
  int feet_sensor_ids[4];        /* contact-sensor ids of ".*foot" in body order */
  int thigh_calf_sensor_ids[8];
  int num_thigh_calf;
  const float* current_air_time;     /* [N,num_sensor_bodies] */
  const float* current_contact_time; /* [N,num_sensor_bodies] */
  const float* last_air_time;        /* [N,num_sensor_bodies] */
  const int64_t* episode_length_buf; /* [N] */
  /* object (NULL for the locomotion task) */
  const float* obj_root_pos_w;   /* [N,3] */
  const float* obj_root_quat_w;  /* [N,4] */
  const float* obj_root_lin_vel_w;   /* [N,3] */
  const float* obj_root_ang_vel_w;   /* [N,3] */
  const float* obj_projected_gravity_b; /* [N,3] */
  const float* obj_last_contact_time;    /* [N] */
  const float* obj_current_contact_time; /* [N] */
  const float* obj_current_air_time;     /* [N] */
  /* ---- term tables ---- */
  int num_reward_terms;
  LtRewardTerm reward_terms[LT_MAX_REWARD_TERMS];
  int num_termination_terms;
  LtTerminationTerm termination_terms[LT_MAX_TERMINATION_TERMS];
  LtGaitParams gait;
  LtGaitState gait_state;
  int any_nonzero_cmd_override;  /* -1: compute torch.any(non_zero_cmd) on the device (rewards.py:190); 0/1: use this */
  int auto_reset;                /* 1: zero gait state + episode sums of done envs after the reward pass (== the
                                    manager .reset(env_ids) calls that follow in ManagerBasedRLEnv.step) */
  /* ---- reward / termination outputs ---- */
  float* reward;                 /* [N]   reward_buf */
  float* step_reward;            /* [N,num_reward_terms]  value/dt  ([IL] RewardManager._step_reward) */
  float* episode_sums;           /* [num_reward_terms,N]  += value  (one contiguous [N] row per term) */
  float* term_raw;               /* [num_reward_terms,N] unweighted term values or NULL (per-term drop-in cache) */
  uint8_t* term_masks;           /* [num_termination_terms,N] or NULL */
  uint8_t* terminated;           /* [N] */
  uint8_t* time_outs;            /* [N] */
  uint8_t* dones;                /* [N] terminated | time_outs */
  float* episode_log_sums;       /* [num_reward_terms+1] or NULL: += episode sums of reset envs, last = count */
  /* ---- observation phase ---- */
  int num_obs_terms;
  LtObsTerm obs_terms[LT_MAX_OBS_TERMS];
  int history_length;            /* 6 */
  uint8_t* obs_fill;             /* [N] in/out or NULL.  1 = history empty: this observation pass writes the new values into
                                    every history slot ([IL] CircularBuffer first push) and clears the flag.  A reward pass with
                                    auto_reset sets it for done envs (or, when fused with the observation pass, applies it). */
  const float* policy_obs_in;    /* [N, D] previous policy observation (history source) */
  float* policy_obs_out;         /* [N, D] may alias policy_obs_in */
  const float* critic_obs_in;
  float* critic_obs_out;
  const float* u_obs;            /* [N, sum(dim)] uniforms for the policy-group noise, or NULL -> Philox */
  const float* u_obj_euler;      /* [N,3] uniforms for the object quaternion noise (observations.py:78) or NULL */
  uint64_t seed, offset;         /* Philox key / counter; `offset` is also the env-step index used by the cross-env any() hand-over */
  const int64_t* offset_base;    /* optional DEVICE counter added to `offset` (lets a captured CUDA graph advance the step) */
  /* object_state_in_robot_frame parameters (observations.py:38-91) */
  float os_n_min[13], os_n_max[13];       /* additive noise per state slot; the 4 quaternion slots are 0 (observations.py:74-77) */
  float os_euler_min[3], os_euler_max[3]; /* euler-angle noise composed onto the quaternion (observations.py:72-73,78-80) */
  float os_scale[13], os_non_contact[13];
  float os_last_contact_thr, os_current_contact_thr;
  int* any_flag_ws;              /* [2] device ints, zero-initialised once; used for the cross-env any() */
  const int32_t* tables;         /* optional DEVICE copy (16-byte aligned) of lt_mdp_build_tables() output for this term
                                    configuration; NULL: every block rebuilds the tables in shared memory */
  /* Optional fused action term (K0 inside K1; act_new == NULL: off).  JointPositionActionPrevPrev.process_actions (locotouch/mdp/actions.py:30-44,
   * = lt_process_actions, bit for bit) is applied to the action-term state of every env before any term of this launch reads it:
   *   prev_prev_raw = prev_raw ; prev_raw = raw ; raw = clamp(act_new, -clip, clip) * raw_scale ; processed = raw * scale + offset.
   * raw_actions / prev_raw_actions above must be the PRE-update state and are rewritten in place (the launch needs LT_PHASE_REWARDS). */
  const float* act_new;          /* [N,J] the policy's actions of this env step */
  float* act_prev_prev_raw;      /* [N,J] or NULL */
  float* act_processed;          /* [N,J] or NULL */
  const float* act_offset;       /* [N,J] or NULL (default joint positions) */
  float act_clip, act_raw_scale, act_scale;
  /* ActionManager.reset(env_ids) inside the launch (IsaacLab's ManagerBasedRLEnv.step order: rewards -> reset -> observations; reference
   * JointPositionActionPrevPrev.reset, locotouch/mdp/actions.py:46-52): with auto_reset and LT_PHASE_REWARDS, raw_actions / prev_raw_actions /
   * act_prev_prev_raw of every env this launch resets are zeroed in place (processed keeps raw * scale + offset of the pre-reset action) and,
   * with LT_PHASE_OBS, the last_action values of its post-reset observation are built from the zeroed row.  0: the action term is left to the
   * caller (the split-phase order compute_rewards -> term.reset(env_ids) -> compute_observations gives the same result). */
  int act_reset_on_done;
  /* Optional fused rollout store (K3 inside K1; store_rewards == NULL: off; needs LT_PHASE_REWARDS).  PPO.process_env_step's time-out bootstrap
   * + the scalar part of RolloutStorage.add_transitions (loco_rl/algorithms/ppo.py:162-165, storage/rollout_storage.py:86-88, = lt_store_step
   * bit for bit): store_rewards[n] = reward[n] + store_gamma * (store_values[n] * time_outs[n]) ; store_dones[n] = dones[n], written straight
   * into the rollout slot of this env step (store_values == NULL: no bootstrap). */
  float* store_rewards;          /* [N] RolloutStorage.rewards[step] */
  uint8_t* store_dones;          /* [N] RolloutStorage.dones[step] or NULL */
  const float* store_values;     /* [N] the critic's values of this step (RolloutStorage.values[step]) or NULL */
  float store_gamma;
} LtMdpArgs;
int lt_mdp_step(const LtMdpArgs* args, void* stream);
/* Launch-constant lookup tables (observation column map, per-value term info, reward kind -> slot); host-side, no CUDA call.
 * They depend on obs_terms / history_length / reward_terms (kinds and zero weights) only.  lt_mdp_tables_len: ints needed
 * (-1 on an invalid table); lt_mdp_build_tables fills `out` (host memory, `len` ints). */
int lt_mdp_tables_len(const LtMdpArgs* args);
int lt_mdp_build_tables(const LtMdpArgs* args, int32_t* out, int len);
/* AdaptiveSymmetricGaitReward.reset(env_ids) (rewards.py:107-114) for a mask of envs; also zeroes episode sums. */
int lt_mdp_reset(const LtGaitState* gait_state, float* episode_sums, int num_reward_terms, const uint8_t* mask,
                 int N, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K8  student batch helpers
 * replaces  locotouch/distill/replay_buffer.py:90-112 (_prepare_padded_sequence) and
 *           locotouch/distill/student.py:131,142-143 (per-element MSE .mean(-1), masked mean) + its backward
 * ------------------------------------------------------------------------------------------------------------------ */
/* Pads B trajectories stored back to back in `flat` [total_steps, D] (trajectory b = rows offsets[b]..offsets[b]+len[b])
 * into out [L_max, B, D] (zero filled) and masks [L_max, B]. */
int lt_pad_trajectories(const float* flat, const int64_t* offsets, const int64_t* lengths, int B, int L_max, int D,
                        float* out, uint8_t* masks, void* stream);
/* loss = sum_{t,b} mask * mean_a (s - t)^2 / sum(mask) ; grad_student = dloss/ds.  out[4] = {loss, mae, count, 0}. */
int64_t lt_masked_mse_workspace_bytes(int64_t rows);
int lt_masked_mse(const float* student, const float* teacher, const uint8_t* masks, int64_t rows, int A,
                  float* grad_student, float* out, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K10  trajectory split / pad / unpad for recurrent mini-batches
 * replaces  loco_rl/loco_rl/utils/utils.py:37-83 (split_and_pad_trajectories, unpad_trajectories) as used by
 *           loco_rl/loco_rl/storage/rollout_storage.py:246-318 (recurrent_mini_batch_generator).
 * Trajectories are numbered env by env in time order; the last step always ends one.  traj_base[n] = exclusive scan over
 * envs of (1 + number of dones before the last step); M = total number of trajectories (the caller sizes the outputs).
 * ------------------------------------------------------------------------------------------------------------------ */
/* (env, first step, length) of every trajectory, from dones [T, N] (uint8). */
int lt_trajectory_index(const uint8_t* dones, const int64_t* traj_base, int32_t* traj_env, int32_t* traj_start,
                        int32_t* traj_len, int T, int N, void* stream);
/* x [T, N, D] -> out [T, M, D] (zero padded) and masks [T, M] (may be NULL). */
int lt_split_pad_trajectories(const float* x, const int32_t* traj_env, const int32_t* traj_start, const int32_t* traj_len,
                              float* out, uint8_t* masks, int T, int N, int D, int M, void* stream);
/* padded [T, M, D] -> out [T, N, D] (every row written exactly once). */
int lt_unpad_trajectories(const float* padded, const int32_t* traj_env, const int32_t* traj_start, const int32_t* traj_len,
                          float* out, int T, int N, int D, int M, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K11  device-side DAgger replay-buffer bookkeeping
 * replaces  locotouch/distill/replay_buffer.py:52-80 (the per-step done handling of collect_data) and the packing of
 *           finished trajectories (:75-80) into the store that lt_pad_trajectories reads (:82-112).
 * ------------------------------------------------------------------------------------------------------------------ */
/* One env step of the collection.  reward_sums += reward; for every done env (in env-index order): append (reward_sums,
 * length) to ep_reward / ep_length and zero the sum; while state[0] < limit also append one (env, first step, length)
 * trajectory record, set start_idx[env] = step_now and add the length to state[0] (the reference's budget test + break).
 * state (device int64[4]): [0] recorded steps, [1] trajectory records written, [2] episodes logged.  always_restart: set
 * start_idx of every done env even when it is not recorded (evaluation).  The output arrays must have room for N more
 * entries.  N <= 65536. */
int lt_dagger_step(const uint8_t* dones, const float* reward, float* reward_sums, int32_t* start_idx, int N, int step_now,
                   int64_t limit, int always_restart, int64_t* state, int32_t* traj_env, int32_t* traj_start,
                   int32_t* traj_len, float* ep_reward, int32_t* ep_length, void* stream);
/* Copies M trajectories out of a step-major buffer x [S, N, D] (trajectory j = rows (traj_start[j] + p, traj_env[j]),
 * p < length) back to back into flat [total_rows, D]; traj_offset[j] = first flat row of trajectory j (ascending). */
int lt_pack_trajectories(const float* x, const int32_t* traj_env, const int32_t* traj_start, const int64_t* traj_offset,
                         int M, int64_t total_rows, int N, int D, float* flat, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K12  fused linear layer: out[M, N] = act(x[M, K] . w[N, K]^T + bias[N]), act = ELU (alpha 1) or identity
 * replaces  nn.Linear (+ nn.ELU) of loco_rl/loco_rl/modules/actor_critic.py:33-56 and models/mlp.py:4-25 (cuBLAS GEMM +
 *           an elementwise ELU launch) with one tcgen05 TF32 GEMM whose epilogue adds the bias and applies the activation.
 * K and N must be multiples of 4 and all pointers 16-byte aligned; LT_ERR_UNSUPPORTED otherwise (or when the library was
 * built without the CUTLASS headers) -- the caller then uses its cuBLAS path.
 * ------------------------------------------------------------------------------------------------------------------ */
const char* lt_gemm_backend(void); /* "stub" when the library was built without the tcgen05 GEMMs (callers then raise / fall back to cuBLAS) */
int64_t lt_linear_bias_act_workspace_bytes(int M, int N, int K);
int lt_linear_bias_act(const float* x, const float* w, const float* bias, float* out, int M, int N, int K, int apply_elu,
                       void* workspace, int64_t workspace_bytes, void* stream);
/* Backward companion: grad_in[M, Kin] = (grad_out[M, Nout] . w[Nout, Kin]) * elu'(act_in[M, Kin]), act_in = the stored
 * post-ELU activation that fed the layer (elu' = act_in > 0 ? 1 : act_in + 1) -- the dgrad GEMM of a layer with the ELU
 * backward of the layer below in its epilogue (replaces a cuBLAS GEMM + the elu_backward pass of autograd).  Same shape /
 * alignment rules and workspace as lt_linear_bias_act. */
int lt_dgrad_act_bwd(const float* grad_out, const float* w, const float* act_in, float* grad_in, int M, int Nout, int Kin,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K15  weight + bias gradient of a Linear layer with in-kernel split-K
 * replaces  autograd's `grad_output.t().mm(input)` and `grad_output.sum(0)` per nn.Linear of
 *           loco_rl/loco_rl/modules/actor_critic.py:33-56 under loss.backward() (loco_rl/loco_rl/algorithms/ppo.py:350), i.e. a
 *           cuBLAS split-K GEMM + its reduction kernel + a column-sum kernel
 * dw[n_out, k_in] (+)= grad_out[B, n_out]^T . act_in[B, k_in], dbias[n_out] (+)= column sums of grad_out (dbias may be NULL):
 * a hand-written tcgen05 (kind::tf32) kernel; both operands are read as they lie (row-major = MN-major, TMA), every CTA
 * accumulates a 128-row slab of dw over its full width (<= 512 columns: the whole TMEM) for a slice of the batch and adds it into
 * dw with red.global.add.v4.f32 (no partial buffers, no second pass); the bias gradient is summed from the grad_out tiles while
 * they sit in shared memory.  zero_first != 0 clears dw (and dbias) on the stream first; with 0 the caller has cleared them (or
 * wants accumulation).  n_out <= 16 (the action-mean / value heads) runs on CUDA cores.
 * Wider layers need n_out % 4 == 0, k_in % 4 == 0 and 16-byte aligned pointers (LT_ERR_UNSUPPORTED otherwise or in a stub build).
 * Summation order over the batch slices is not fixed (fp32 atomics): results are reproducible to rounding, not bit for bit. */
int lt_wgrad_splitk(const float* grad_out, const float* act_in, float* dw, float* dbias, int B, int n_out, int k_in, int zero_first, void* stream);
/* Two layers of identical shape (layer i of the actor and of the critic) in ONE launch: the CTAs are divided between the two problems, the
 * fixed cost of a launch is paid once.  Outputs are accumulated into (the caller cleared them); dbias0 / dbias1 both NULL or both set;
 * n_out > 16 only.  LT_ERR_UNSUPPORTED under the alignment rules of lt_wgrad_splitk. */
int lt_wgrad_splitk_pair(const float* grad_out0, const float* act_in0, float* dw0, float* dbias0, const float* grad_out1, const float* act_in1,
                         float* dw1, float* dbias1, int B, int n_out, int k_in, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K19  the three hidden layers (k0 -> 512 -> 256 -> 128, bias + ELU each) of up to two MLPs in ONE persistent tcgen05 kernel
 * replaces  the three nn.Linear + nn.ELU pairs of actor and critic (loco_rl/loco_rl/modules/actor_critic.py:33-56) as `act` /
 *           `evaluate` run them per env step (actor_critic.py:105-131 from algorithms/ppo.py:129-141) and per mini-batch
 *           (algorithms/ppo.py:264-281): six K12 launches (reference: six cuBLAS GEMMs + six ELU launches)
 * A CTA carries a 128-row slab of one network through all three layers: x slab resident in shared memory, weight tiles streamed
 * by TMA, accumulators in TMEM, bias + ELU written back to TMEM in place where they are the A operand of the next layer
 * (kind::tf32, like the reference's TF32 training configuration).  h3 [B,128] is always written; h1 [B,512] / h2 [B,256] only when
 * not NULL (the training pass keeps them for the backward).  Both nets must share k0; k0 % 4 == 0, k0 <= 352 (row pitch of x and
 * w1), all pointers 16-byte aligned; LT_ERR_UNSUPPORTED otherwise (callers keep the per-layer K12 path).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct LtMlp3Net {
  const float* x;              /* [B, k0] */
  int k0;
  const float *w1, *b1;        /* [512, k0], [512] */
  const float *w2, *b2;        /* [256, 512], [256] */
  const float *w3, *b3;        /* [128, 256], [128] */
  float *h1, *h2, *h3;         /* post-ELU activations; h1 / h2 may be NULL */
} LtMlp3Net;
int lt_mlp3_forward(const LtMlp3Net* nets, int n_nets, int B, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K3b  output heads of both MLPs + action sampling + log-prob, one launch per env step of the rollout
 * replaces  the head nn.Linear of actor and critic + Normal.sample() + log_prob().sum(-1) of
 *           loco_rl/loco_rl/modules/actor_critic.py:105-131 as PPO.act drives them (algorithms/ppo.py:129-141): a cuBLAS GEMM, a GEMV,
 *           lt_act_sample and the value copy of round 1
 * h_* [N,H] post-ELU activations of the last hidden layers (H % 128 == 0, H <= 256; h_critic may be NULL: actor only), heads as in
 * K16; sigma [A] (A % 4 == 0, A <= 16); eps [N,A] explicit standard-normal draws or NULL = Philox (the stream of lt_act_sample:
 * key (seed, offset + *offset_base), counter (env, chunk)).  Outputs: actions / mu_out / sigma_out [N,A], logp [N], values [N].
 * ------------------------------------------------------------------------------------------------------------------ */
int lt_act_heads(const float* h_actor, const float* h_critic, const float* w_actor, const float* b_actor, const float* w_critic,
                 const float* b_critic, const float* sigma, const float* eps, float* actions, float* logp, float* mu_out,
                 float* sigma_out, float* values, int N, int A, int H, uint64_t seed, uint64_t offset, const int64_t* offset_base,
                 void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K17  tactile pre-encoder of the CNN-RNN student, forward, one kernel per batch of frames
 * replaces  loco_rl/loco_rl/models/cnn_2d.py:16-131 (CNN2dHead.forward = 3 x [Conv2d, ReLU(, MaxPool2d)] + flatten + MLP head) as
 *           built by loco_rl/loco_rl/models/model_generation.py:16-20 from locotouch/config/locotouch/agents/distillation_cfg.py:78-85
 *           and called by locotouch/distill/student.py:88-103 (8 cuDNN / cuBLAS / ATen launches per call in the reference)
 * Geometry taken: image (2,17,13), channels (24,24,24), kernels (4,3,2), MaxPool2d(2) after conv1 only, ReLU, no padding, head =
 * one Linear 192 -> embedding_dim <= 64 (LT_ERR_UNSUPPORTED for anything else: the torch modules stay in charge).
 * Input: `image` [M, 442] fp32 (channel-major like the observation) or `packed` [M, packed_words] uint32 (bit t%32 of word t/32 =
 * taxel t of the 17 x 13 grid, both channels equal -- the bitmap K2 emits); exactly one of the two may be NULL.
 * Weights are the nn.Module tensors as they lie: w1 [24,2,4,4], w2 [24,24,3,3], w3 [24,24,2,2], wh [E,192], biases [24]/[E].
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int M;                       /* frames */
  int in_channels, height, width;
  int channels[3], kernel_sizes[3], pool[3];
  int embedding_dim;
  const float* image;
  const uint32_t* packed;
  int packed_words;
  const float *w1, *b1, *w2, *b2, *w3, *b3, *wh, *bh;
  float* out;                  /* [M, embedding_dim] */
} LtStudentCnnArgs;
int lt_student_cnn_forward(const LtStudentCnnArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K18  ContactSensor bookkeeping: net-force history ring + air / contact-time state machine, one launch per sensor and step
 * replaces  [IL] isaaclab.sensors.ContactSensor._update_buffers_impl / reset (IsaacLab 2.2.1; configured by the reference in
 *           locotouch/config/locotouch/locomotion_base_env_cfg.py:35-39,358-359: history_length 3, track_air_time) whose outputs
 *           the reference terms read (locotouch/mdp/rewards.py:116-156,596-604, observations.py:60-66)
 * forces [N,Bd,3] = this step's net contact forces; net_forces_w (may alias forces, may be NULL) receives them; history
 * [N,H,Bd,3] is shifted by one slot and gets them in slot 0 (NULL: no history); the four timers [N,Bd] follow
 *   is_contact = |F| > force_threshold; last_air = (cur_air > 0 & contact) ? cur_air + dt : last_air; cur_air = contact ? 0 : cur_air + dt;
 *   last_contact = (cur_contact > 0 & !contact) ? cur_contact + dt : last_contact; cur_contact = contact ? cur_contact + dt : 0
 * (all four NULL: forces / history only).  dt_per_env [N] overrides the scalar dt when not NULL.  Envs with reset_mask[n] != 0 get
 * ContactSensor.reset instead: forces, history and timers cleared.
 * ------------------------------------------------------------------------------------------------------------------ */
int lt_contact_sensor_update(const float* forces, float* net_forces_w, float* net_forces_w_history, int history_length, int N, int num_bodies,
                             float* current_air_time, float* last_air_time, float* current_contact_time, float* last_contact_time,
                             const float* dt_per_env, float dt, float force_threshold, const uint8_t* reset_mask, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K14  gradient all-reduce folded into the optimizer step (env-sharded data parallelism)
 * replaces  the NCCL all-reduce of the flat PPO gradient + nn.utils.clip_grad_norm_ + Adam.step per mini-batch
 *           (loco_rl/loco_rl/algorithms/ppo.py:350-353 run under one process per GPU)
 * peer_grads[r] = device-visible address of rank r's flat gradient buffer [n + tail] (NVLink peer mapping / symmetric memory;
 * entry `rank` is the caller's own buffer).  The caller separates the ranks' writes from these reads with a cross-GPU barrier
 * (and the reads from the next writes with another).  grad_sum [n + tail] (local) receives sum_r peer_grads[r] added in rank
 * order -- bit-identical on every rank --; the norm, clip and Adam update then run on grad_sum * grad_scale exactly as in
 * lt_clip_adam (same workspace).  `tail` extra floats behind the gradients are summed but not part of the norm; with
 * desired_kl > 0 the first of them is the per-rank KL mean and *lr is adapted from sum * kl_scale (ppo.py:275-281) before the step.
 * n must be a multiple of 4; world <= LT_MAX_PEERS.
 * ------------------------------------------------------------------------------------------------------------------ */
#define LT_MAX_PEERS 16
int lt_peer_sum_clip_adam(float* params, const float* const* peer_grads, int world, float* grad_sum, int tail, float* exp_avg,
                          float* exp_avg_sq, int64_t n, float* lr, float* step_inout, float max_grad_norm, double beta1,
                          double beta2, float eps, float weight_decay, float grad_scale, float desired_kl, float kl_scale,
                          float* grad_norm_out, void* workspace, int64_t workspace_bytes, void* stream);

/* K14, two-shot variant for larger worlds (W - 1 remote buffers per rank become 2 (W - 1) / W):
 * lt_peer_reduce_scatter: rank `rank` adds slice `rank` (ceil(n / 4 / world) float4 each, the last slice takes the rest) of all W
 * buffers in rank order and writes the sum over that slice of ITS OWN buffer peer_grads[rank]; then a cross-GPU barrier; then
 * lt_peer_gather_clip_adam (same arguments as lt_peer_sum_clip_adam) reads slice q from rank q instead of summing W buffers.  The tail
 * statistics are still summed directly.  Results are bit-identical to the one-shot exchange (same summation order per element).
 * LT_ERR_UNSUPPORTED when n exceeds what the one-launch optimizer kernel holds (use the one-shot call). */
int lt_peer_reduce_scatter(const float* const* peer_grads, int world, int rank, int64_t n, void* stream);
int lt_peer_gather_clip_adam(float* params, const float* const* peer_grads, int world, float* grad_sum, int tail, float* exp_avg,
                          float* exp_avg_sq, int64_t n, float* lr, float* step_inout, float max_grad_norm, double beta1,
                          double beta2, float eps, float weight_decay, float grad_scale, float desired_kl, float kl_scale,
                          float* grad_norm_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K13  velocity command term + reward-driven velocity curriculum, device resident
 * replaces  locotouch/mdp/commands.py:379-576  UniformVelocityCommandGaitLoggingMultiSampling (reset / compute / set_ranges,
 *           over IsaacLab's CommandTerm + UniformVelocityCommand) and
 *           locotouch/mdp/curriculums.py:184-274  ModifyVelCommandsRangeBasedonReward.__call__
 * The reference runs both as per-reset host logic (nonzero, multinomial, boolean-index writes, torch.all / torch.mean read on
 * the host, ranges as Python tuples).  Here the ranges live in one device block that both entry points share, env ids are a
 * device mask and no host read is needed between a reset and the next command.
 * ------------------------------------------------------------------------------------------------------------------ */
/* Scalar state of the command term (cfg.ranges, cfg.previous_ranges, *_equal_ranges, initial_zero_command_steps,
 * cfg.rel_standing_envs) and of the curriculum term (forward bins, success counters).  Doubles: the reference keeps these as
 * Python floats and decides *_equal_ranges by tuple equality.  Index 0 lin_vel_x, 1 lin_vel_y, 2 ang_vel_z; [lo, hi]. */
typedef struct LtCommandRanges {
  double ranges[3][2];
  double previous[3][2];
  int32_t equal[3];
  int32_t initial_zero_command_steps;
  double rel_standing_envs;
  int32_t final_initial_zero_command_steps;
  int32_t reserved;
  double final_rel_standing_envs;
  int32_t lin_forward_bins, ang_forward_bins, success_repeat_times_lin, success_repeat_times_ang;
} LtCommandRanges;

#define LT_CMD_RESET 1   /* [IL] CommandTerm.reset(env_ids), env ids given as reset_mask */
#define LT_CMD_COMPUTE 2 /* [IL] CommandTerm.compute(dt) */
/* rows of the [LT_CMD_NUM_METRICS, N] metrics block (= the reference's metrics dict of [N] tensors) */
enum LtCommandMetric {
  LT_CMD_M_ERROR_VEL_XY = 0, LT_CMD_M_ERROR_VEL_YAW, LT_CMD_M_FOOT_AIR_TIME_VAR, LT_CMD_M_FOOT_STEP_FREQ, LT_CMD_M_PAIR1_STEP_FREQ,
  LT_CMD_M_PAIR2_STEP_FREQ, LT_CMD_M_STEP_AIR_TIME, LT_CMD_M_PAIR1_AIR_TIME, LT_CMD_M_PAIR2_AIR_TIME, LT_CMD_M_LIN_VEL_X,
  LT_CMD_M_LIN_VEL_Y, LT_CMD_M_ANG_VEL_Z, LT_CMD_M_ZERO_STEPS, LT_CMD_M_REL_STANDING, LT_CMD_NUM_METRICS
};

typedef struct LtCommandArgs {
  int32_t N;
  int32_t phases;                  /* LT_CMD_RESET | LT_CMD_COMPUTE (reset runs first) */
  float dt;                        /* env.step_dt */
  float resampling_time_lo, resampling_time_hi; /* cfg.resampling_time_range */
  float bin_c0, bin_c1;            /* normalised cumulative [p, 1 - p] of sampling_probs (commands.py:448), fp32 */
  int32_t binary_maximal_command;
  const LtCommandRanges* ranges;   /* device */
  /* term state, read-write */
  float* vel_command_b;            /* [N,3] the command tensor lt_mdp_step reads */
  float* vel_command_b_buffer;     /* [N,3] */
  float* time_left;                /* [N] */
  int64_t* command_counter;        /* [N] */
  uint8_t* is_standing_env;        /* [N] */
  float* metrics;                  /* [LT_CMD_NUM_METRICS, N] */
  float* metric_scalars;           /* [LT_CMD_NUM_METRICS] launch-wide values of rows 3..13 (written by COMPUTE) */
  /* inputs */
  const uint8_t* reset_mask;       /* [N], RESET */
  double* reset_extras;            /* [LT_CMD_NUM_METRICS + 1], RESET: += sum of every metric row over the reset envs, last =
                                      their count (the caller zeroes it; mean = sum / count is what the reference logs) */
  const int64_t* episode_length_buf;
  const float* root_lin_vel_b;     /* [N,3], COMPUTE */
  const float* root_ang_vel_b;     /* [N,3], COMPUTE */
  const float* last_air_time;      /* [N, num_sensor_bodies], COMPUTE */
  int32_t num_sensor_bodies;
  int32_t feet_ids[4];             /* sensor_cfg.body_ids */
  const float* gait_valid_last_air_time; /* [N,4] state of the gait reward term (rewards.py:99) or NULL; 16-byte aligned */
  /* randomness: explicit uniforms [N,8] (slot 0 time_left, 1-3 x / y / yaw value, 4-6 bin, 7 standing) for ONE phase, or NULL:
   * Philox4x32-10 keyed by (seed, offset + *offset_base), counter (env, phase) */
  const float* u;
  uint64_t seed, offset;
  const int64_t* offset_base;
  void* workspace;                 /* COMPUTE: lt_command_workspace_bytes(N), zero-initialised once */
  int64_t workspace_bytes;
} LtCommandArgs;

int64_t lt_command_workspace_bytes(int N);
int lt_command_step(const LtCommandArgs* args, void* stream);

typedef struct LtVelCurriculumArgs {
  int32_t N;
  int32_t repeat_times_lin, repeat_times_ang, max_distance_bins;
  LtCommandRanges* ranges;            /* device, read-write */
  const uint8_t* reset_mask;          /* [N] env_ids of the call */
  const int64_t* episode_length_buf;  /* [N] */
  const float* episode_sums_lin;      /* [N] RewardManager._episode_sums[reward_name_lin] */
  const float* episode_sums_ang;      /* [N] */
  uint8_t* env_reseted_lin; float* episode_length_buf_lin; float* episode_reward_sum_lin; /* term state [N] */
  uint8_t* env_reseted_ang; float* episode_length_buf_ang; float* episode_reward_sum_ang;
  double command_maximum_ranges[3];
  double expansion[3];                /* (maximum - initial upper bound) / curriculum_bins, curriculums.py:191-193 */
  double reset_envs_episode_length;   /* cfg value * max_episode_length_s (:194) */
  double reward_threshold_lin, reward_threshold_ang; /* :199-200 */
} LtVelCurriculumArgs;

/* One block; N <= 1 << 20. */
int lt_vel_curriculum(const LtVelCurriculumArgs* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LOCOTOUCH_B200_H */
