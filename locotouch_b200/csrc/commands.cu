// K13: velocity command term + reward-driven velocity curriculum on the device.
// Replaces reference locotouch/mdp/commands.py:379-576 (UniformVelocityCommandGaitLoggingMultiSampling over [IL] CommandTerm /
// UniformVelocityCommand: per-reset and per-step host logic with nonzero / multinomial / boolean-index writes and .item() reads)
// and locotouch/mdp/curriculums.py:184-274 (ModifyVelCommandsRangeBasedonReward: torch.all / torch.mean decisions read on the
// host at every reset, Python tuples of ranges).
//
//   lt_command_step   LT_CMD_RESET   [IL] CommandTerm.reset(env_ids) with the env ids as a device mask: sums of the 14 metric rows
//                                    over the reset envs (the logged means), rows cleared, counters zeroed, time_left + command
//                                    resampled (multi-sampling bins / uniform / binary-maximal), buffer updated, zero-command rule.
//                     LT_CMD_COMPUTE [IL] CommandTerm.compute(dt): metrics (tracking errors, foot air-time variance, the masked
//                                    means of the gait term's valid_last_air_time as a grid reduction), time_left -= dt, resampling
//                                    where it expired, zero-command / recover rule, standing envs.  A second small kernel broadcasts
//                                    the launch-wide scalars into their [N] metric rows (the reference stores them per env).
//   lt_vel_curriculum one block: masked capture of episode length / reward sums for the reset envs, torch.all / torch.mean decisions,
//                     range expansion (np.clip in double), set_ranges bookkeeping (previous ranges, *_equal_ranges, switch to the
//                     final zero-command steps / standing share) -- all in the device-resident LtCommandRanges block that
//                     lt_command_step reads, so no host read sits between a reset and the next command.
// Ranges are doubles on the device because the reference keeps them as Python floats (tuple equality decides *_equal_ranges).
// Built with -fmad=false (metric arithmetic in torch's two-rounding order); the uniform transform is an explicit fmaf, as in torch.
#include "lt_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kCurThreads = 1024;

struct GaitPartial { double n, s01, s23; };

__device__ __forceinline__ float uniform_in(float u, double lo, double hi) {
  const float lo32 = (float)lo, hi32 = (float)hi;
  return __fmaf_rn(u, hi32 - lo32, lo32);  // torch's uniform_ transform is one fused multiply-add (tests/test_commands.py pins it)
}

// commands.py:517-559 for one env.  u: 8 uniforms (slot 0 time_left, 1-3 value, 4-6 bin, 7 standing).
__device__ __forceinline__ void resample_env(const LtCommandArgs& A, const LtCommandRanges& R, int n, const float u[8], float cmd[3]) {
  A.time_left[n] = uniform_in(u[0], A.resampling_time_lo, A.resampling_time_hi);  // [IL] CommandTerm._resample
  if (A.binary_maximal_command) {  // :518-521, combos ordered (i, j, k) over (-1, 1)
    int idx = (int)(u[1] * 8.f);
    idx = idx > 7 ? 7 : idx;
    cmd[0] = ((idx & 4) ? 1.f : -1.f) * (float)R.ranges[0][1];
    cmd[1] = ((idx & 2) ? 1.f : -1.f) * (float)R.ranges[1][1];
    cmd[2] = ((idx & 1) ? 1.f : -1.f) * (float)R.ranges[2][1];
  } else {
    const bool all_equal = R.equal[0] && R.equal[1] && R.equal[2];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double lo = R.ranges[d][0], hi = R.ranges[d][1];
      if (!all_equal && !R.equal[d]) {  // :528-534 multinomial over [new low | old | new high], then uniform inside the bin
        const float ub = u[4 + d];
        const int bin = (ub >= A.bin_c0 ? 1 : 0) + (ub >= A.bin_c1 ? 1 : 0);
        // the reference keeps the bins in an fp32 tensor: round first
        const float b_lo = bin == 0 ? (float)R.ranges[d][0] : (bin == 1 ? (float)R.previous[d][0] : (float)R.previous[d][1]);
        const float b_hi = bin == 0 ? (float)R.previous[d][0] : (bin == 1 ? (float)R.previous[d][1] : (float)R.ranges[d][1]);
        lo = b_lo;
        hi = b_hi;
      }
      cmd[d] = uniform_in(u[1 + d], lo, hi);
    }
    A.is_standing_env[n] = uniform_in(u[7], 0.0, 1.0) <= (float)R.rel_standing_envs ? 1 : 0;  // :555 / [IL]
  }
  A.vel_command_b_buffer[3 * n] = cmd[0];  // :558
  A.vel_command_b_buffer[3 * n + 1] = cmd[1];
  A.vel_command_b_buffer[3 * n + 2] = cmd[2];
  A.command_counter[n] += 1;  // [IL] CommandTerm._resample
}

__device__ __forceinline__ void draw8(const LtCommandArgs& A, int n, int phase, float u[8]) {
  if (A.u) {
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] = A.u[(size_t)n * 8 + k];
  } else {
    const uint64_t off = A.offset + (A.offset_base ? (uint64_t)*A.offset_base : 0ull);
    const uint4 a = lt::Philox::gen(A.seed, off, (uint32_t)n, 0x4000u + 2u * phase);
    const uint4 b = lt::Philox::gen(A.seed, off, (uint32_t)n, 0x4001u + 2u * phase);
    u[0] = lt::Philox::u01(a.x); u[1] = lt::Philox::u01(a.y); u[2] = lt::Philox::u01(a.z); u[3] = lt::Philox::u01(a.w);
    u[4] = lt::Philox::u01(b.x); u[5] = lt::Philox::u01(b.y); u[6] = lt::Philox::u01(b.z); u[7] = lt::Philox::u01(b.w);
  }
}

// ws layout (doubles): [0..2] gait partial totals (n, s01, s23), [3] block counter (as unsigned), then per-block partials
__global__ void __launch_bounds__(kThreads) command_step_kernel(const LtCommandArgs A, int phase) {
  __shared__ double s_red[kThreads / 32];
  __shared__ LtCommandRanges s_R;
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  if (tid < (int)(sizeof(LtCommandRanges) / 4)) reinterpret_cast<int*>(&s_R)[tid] = reinterpret_cast<const int*>(A.ranges)[tid];
  __syncthreads();
  const LtCommandRanges& R = s_R;
  const int N = A.N;
  const int n = blockIdx.x * kThreads + tid;
  const bool live = n < N;
  const int izcs = R.initial_zero_command_steps;

  if (phase == LT_CMD_RESET) {
    // ---- [IL] CommandTerm.reset: mean of every metric row over the reset envs (sums + count here), rows cleared for them
    const bool rs = live && A.reset_mask[n];
    // only warps that hold a reset env reduce (a few per cent of the envs reset in a typical step): warp sums, one atomic per
    // row and warp
    if (__ballot_sync(LT_FULL_MASK, rs)) {
      const int lane = tid & 31;
#pragma unroll 1
      for (int m = 0; m < LT_CMD_NUM_METRICS; ++m) {
        double v = 0.0;
        if (rs) {
          v = A.metrics[(size_t)m * N + n];
          A.metrics[(size_t)m * N + n] = 0.f;
        }
        v = lt::warp_sum(v);
        if (lane == 0 && v != 0.0) atomicAdd(&A.reset_extras[m], v);
      }
      const double c = lt::warp_sum(rs ? 1.0 : 0.0);
      if (lane == 0) atomicAdd(&A.reset_extras[LT_CMD_NUM_METRICS], c);
    }
    if (!live) return;
    float cmd[3] = {A.vel_command_b[3 * n], A.vel_command_b[3 * n + 1], A.vel_command_b[3 * n + 2]};
    if (rs) {
      A.command_counter[n] = 0;
      float u[8];
      draw8(A, n, 0, u);
      resample_env(A, R, n, u, cmd);
    }
    // commands.py:559 -> :566-570: zero command while the episode is younger than initial_zero_command_steps (buffer * 0.0 keeps
    // the sign of the buffered value, like the reference)
    if (A.episode_length_buf[n] < izcs) {
#pragma unroll
      for (int d = 0; d < 3; ++d) cmd[d] = A.vel_command_b_buffer[3 * n + d] * 0.0f;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) A.vel_command_b[3 * n + d] = cmd[d];
    return;
  }

  // ------------------------------------------------------------------------------------------------ LT_CMD_COMPUTE
  GaitPartial gp = {0.0, 0.0, 0.0};
  if (live) {
    float cmd[3] = {A.vel_command_b[3 * n], A.vel_command_b[3 * n + 1], A.vel_command_b[3 * n + 2]};
    // ---- _update_metrics (commands.py:393-397): tracking errors against the command as it stands BEFORE this step's update
    const float dx = cmd[0] - A.root_lin_vel_b[3 * n], dy = cmd[1] - A.root_lin_vel_b[3 * n + 1];
    A.metrics[(size_t)LT_CMD_M_ERROR_VEL_XY * N + n] = sqrtf(dx * dx + dy * dy);
    A.metrics[(size_t)LT_CMD_M_ERROR_VEL_YAW * N + n] = fabsf(cmd[2] - A.root_ang_vel_b[3 * n + 2]);
    float la[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) la[k] = A.last_air_time[(size_t)n * A.num_sensor_bodies + A.feet_ids[k]];
    const float mean = (((la[0] + la[1]) + la[2]) + la[3]) / 4.f;
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) var += (la[k] - mean) * (la[k] - mean);
    A.metrics[(size_t)LT_CMD_M_FOOT_AIR_TIME_VAR * N + n] = var / 3.f;  // torch.var: unbiased
    if (A.gait_valid_last_air_time) {  // :400-403 envs whose four feet all hold a valid last air time
      const float4 v = *reinterpret_cast<const float4*>(A.gait_valid_last_air_time + 4 * (size_t)n);
      if (v.x > 1.0e-6f && v.y > 1.0e-6f && v.z > 1.0e-6f && v.w > 1.0e-6f) {
        gp.n = 1.0;
        gp.s01 = (double)v.x + (double)v.y;
        gp.s23 = (double)v.z + (double)v.w;
      }
    }
    // ---- [IL] CommandTerm.compute: timer, resampling where it expired
    const float tl = A.time_left[n] - A.dt;
    A.time_left[n] = tl;
    if (tl <= 0.f) {
      float u[8];
      draw8(A, n, 1, u);
      resample_env(A, R, n, u, cmd);
    }
    // ---- _update_command (:561-576 + [IL] standing envs)
    const long long len = A.episode_length_buf[n];
    if (len < izcs) {
#pragma unroll
      for (int d = 0; d < 3; ++d) cmd[d] = A.vel_command_b_buffer[3 * n + d] * 0.0f;
    }
    if (len == izcs) {
#pragma unroll
      for (int d = 0; d < 3; ++d) cmd[d] = A.vel_command_b_buffer[3 * n + d];
    }
    if (A.is_standing_env[n]) cmd[0] = cmd[1] = cmd[2] = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) A.vel_command_b[3 * n + d] = cmd[d];
  }
  // ---- grid reduction of the gait statistics: block partials, the last block folds them in block order (deterministic) and
  // writes the launch-wide scalars (commands.py:404-418, 507-513)
  double* ws = reinterpret_cast<double*>(A.workspace);
  double* part = ws + 4;
  const double bn = lt::block_sum(gp.n, s_red), b01 = lt::block_sum(gp.s01, s_red), b23 = lt::block_sum(gp.s23, s_red);
  if (tid == 0) {
    part[3 * blockIdx.x] = bn;
    part[3 * blockIdx.x + 1] = b01;
    part[3 * blockIdx.x + 2] = b23;
    __threadfence();
    const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(ws + 3), 1u);
    s_last = done == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || tid != 0) return;
  __threadfence();
  double cnt = 0.0, s01 = 0.0, s23 = 0.0;
  for (unsigned b = 0; b < gridDim.x; ++b) {
    cnt += part[3 * b];
    s01 += part[3 * b + 1];
    s23 += part[3 * b + 2];
  }
  *reinterpret_cast<unsigned*>(ws + 3) = 0u;  // ready for the next launch
  float* sc = A.metric_scalars;
  if (A.gait_valid_last_air_time) {
    // torch.mean over an empty selection is nan and `nan > 0` is false: every value falls back to 0 (:405-418)
    const float avg = cnt > 0.0 ? (float)((s01 + s23) / (4.0 * cnt)) : 0.f;
    const float p1 = cnt > 0.0 ? (float)(s01 / (2.0 * cnt)) : 0.f, p2 = cnt > 0.0 ? (float)(s23 / (2.0 * cnt)) : 0.f;
    sc[LT_CMD_M_FOOT_STEP_FREQ] = avg > 0.f ? 1.0f / avg / 2.0f : 0.f;
    sc[LT_CMD_M_PAIR1_STEP_FREQ] = p1 > 0.f ? 1.0f / p1 / 2.0f : 0.f;
    sc[LT_CMD_M_PAIR2_STEP_FREQ] = p2 > 0.f ? 1.0f / p2 / 2.0f : 0.f;
    sc[LT_CMD_M_STEP_AIR_TIME] = avg > 0.f ? avg : 0.f;
    sc[LT_CMD_M_PAIR1_AIR_TIME] = p1 > 0.f ? p1 : 0.f;
    sc[LT_CMD_M_PAIR2_AIR_TIME] = p2 > 0.f ? p2 : 0.f;
  }
  sc[LT_CMD_M_LIN_VEL_X] = (float)R.ranges[0][1];
  sc[LT_CMD_M_LIN_VEL_Y] = (float)R.ranges[1][1];
  sc[LT_CMD_M_ANG_VEL_Z] = (float)R.ranges[2][1];
  sc[LT_CMD_M_ZERO_STEPS] = (float)R.initial_zero_command_steps;
  sc[LT_CMD_M_REL_STANDING] = (float)R.rel_standing_envs;
}

// the launch-wide scalars go into their [N] rows (the reference assigns `metrics[name][:] = value`)
__global__ void __launch_bounds__(kThreads) command_broadcast_kernel(float* __restrict__ metrics, const float* __restrict__ scalars, int N,
                                                                    int with_gait) {
  const int n = blockIdx.x * kThreads + threadIdx.x;
  if (n >= N) return;
  for (int m = with_gait ? LT_CMD_M_FOOT_STEP_FREQ : LT_CMD_M_LIN_VEL_X; m < LT_CMD_NUM_METRICS; ++m) metrics[(size_t)m * N + n] = scalars[m];
}

// ----------------------------------------------------------------------------------------------- curriculums.py:184-274
__device__ void set_range(LtCommandRanges* R, int d, double lo, double hi) {  // commands.py:471-496 for one dimension
  R->previous[d][0] = R->ranges[d][0];
  R->previous[d][1] = R->ranges[d][1];
  R->ranges[d][0] = lo;
  R->ranges[d][1] = hi;
  R->equal[d] = (R->previous[d][0] == lo && R->previous[d][1] == hi) ? 1 : 0;
}
__device__ void after_set_ranges(LtCommandRanges* R) {  // commands.py:497-505
  if (R->equal[0] && R->equal[1] && R->equal[2]) {
    R->initial_zero_command_steps = R->final_initial_zero_command_steps;
    R->rel_standing_envs = R->final_rel_standing_envs;
  }
}
__device__ __forceinline__ double clip(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

__global__ void __launch_bounds__(kCurThreads) vel_curriculum_kernel(const LtVelCurriculumArgs A) {
  __shared__ double s_red[kCurThreads / 32];
  __shared__ int s_flag;
  const int tid = threadIdx.x, N = A.N;
  LtCommandRanges* R = A.ranges;
  for (int branch = 0; branch < 2; ++branch) {  // 0: linear velocities (:229-251), 1: yaw rate (:252-274), sequentially
    __syncthreads();  // thread 0's writes of the previous branch are visible to the condition below
    bool active;
    if (branch == 0)
      active = (R->ranges[0][1] != A.command_maximum_ranges[0] || !R->equal[0] || R->ranges[1][1] != A.command_maximum_ranges[1] || !R->equal[1]) &&
               R->lin_forward_bins - R->ang_forward_bins <= A.max_distance_bins;
    else
      active = (R->ranges[2][1] != A.command_maximum_ranges[2] || !R->equal[2]) && R->ang_forward_bins - R->lin_forward_bins <= A.max_distance_bins;
    __syncthreads();  // everyone has read R before thread 0 may modify it
    if (!active) continue;
    uint8_t* reseted = branch ? A.env_reseted_ang : A.env_reseted_lin;
    float* len_buf = branch ? A.episode_length_buf_ang : A.episode_length_buf_lin;
    float* sum_buf = branch ? A.episode_reward_sum_ang : A.episode_reward_sum_lin;
    const float* sums = branch ? A.episode_sums_ang : A.episode_sums_lin;
    double len_sum = 0.0, rew_sum = 0.0;
    int all = 1;
#pragma unroll 4
    for (int n = tid; n < N; n += kCurThreads) {  // every load is unconditional and independent: one memory round trip per batch
      const bool rs = A.reset_mask[n] != 0;
      const bool was = reseted[n] != 0;
      const float len_old = len_buf[n], sum_old = sum_buf[n];
      const float len_new = (float)A.episode_length_buf[n], sum_new = sums[n];
      if (rs) {
        reseted[n] = 1;
        len_buf[n] = len_new;
        sum_buf[n] = sum_new;
      }
      all &= (rs || was) ? 1 : 0;
      len_sum += rs ? len_new : len_old;
      rew_sum += rs ? sum_new : sum_old;
    }
    all = __syncthreads_and(all);
    len_sum = lt::block_sum(len_sum, s_red);
    rew_sum = lt::block_sum(rew_sum, s_red);
    if (tid == 0) {
      // torch.mean of an fp32 tensor compared with a Python float: the comparison runs in fp32
      const float thr = (float)(branch ? A.reward_threshold_ang : A.reward_threshold_lin);
      const bool ok = all && (float)(len_sum / N) > (float)A.reset_envs_episode_length && (float)(rew_sum / N) > thr;
      s_flag = ok;
      if (ok) {
        int* success = branch ? &R->success_repeat_times_ang : &R->success_repeat_times_lin;
        *success += 1;
        if (*success == (branch ? A.repeat_times_ang : A.repeat_times_lin)) {
          if (branch == 0) {
            const double lx = clip(R->ranges[0][0] - A.expansion[0], -A.command_maximum_ranges[0], 0.0);
            const double ly = clip(R->ranges[1][0] - A.expansion[1], -A.command_maximum_ranges[1], 0.0);
            set_range(R, 0, lx, -lx);
            set_range(R, 1, ly, -ly);
            R->lin_forward_bins += 1;
          } else {
            const double lz = clip(R->ranges[2][0] - A.expansion[2], -A.command_maximum_ranges[2], 0.0);
            set_range(R, 2, lz, -lz);
            R->ang_forward_bins += 1;
          }
          after_set_ranges(R);
          *success = 0;
        }
      }
    }
    __syncthreads();
    if (s_flag) {
      for (int n = tid; n < N; n += kCurThreads) {
        reseted[n] = 0;
        len_buf[n] = 0.f;
        sum_buf[n] = 0.f;
      }
    }
  }
}

}  // namespace

extern "C" int64_t lt_command_workspace_bytes(int N) {
  if (N <= 0) return -1;
  return (int64_t)(4 + 3 * lt::ceil_div(N, kThreads)) * 8;
}

extern "C" int lt_command_step(const LtCommandArgs* a, void* stream) {
  if (!a || a->N <= 0 || !a->ranges || !a->vel_command_b || !a->vel_command_b_buffer || !a->time_left || !a->command_counter ||
      !a->is_standing_env || !a->episode_length_buf || !a->metrics || !a->metric_scalars)
    return LT_ERR_INVALID_ARG;
  if ((a->phases & ~(LT_CMD_RESET | LT_CMD_COMPUTE)) || !a->phases) return LT_ERR_INVALID_ARG;
  if ((a->phases & LT_CMD_RESET) && (!a->reset_mask || !a->reset_extras)) return LT_ERR_INVALID_ARG;
  if ((a->phases & LT_CMD_COMPUTE) && (!a->root_lin_vel_b || !a->root_ang_vel_b || !a->last_air_time || !a->workspace || a->num_sensor_bodies <= 0))
    return LT_ERR_INVALID_ARG;
  if ((a->phases & LT_CMD_COMPUTE) && a->workspace_bytes < lt_command_workspace_bytes(a->N)) return LT_ERR_WORKSPACE;
  if (a->phases == (LT_CMD_RESET | LT_CMD_COMPUTE) && a->u) return LT_ERR_INVALID_ARG;  // explicit uniforms belong to one phase
  if (a->gait_valid_last_air_time && ((uintptr_t)a->gait_valid_last_air_time & 15)) return LT_ERR_INVALID_ARG;
  const unsigned blocks = (unsigned)lt::ceil_div(a->N, kThreads);
  cudaStream_t s = (cudaStream_t)stream;
  if (a->phases & LT_CMD_RESET) command_step_kernel<<<blocks, kThreads, 0, s>>>(*a, LT_CMD_RESET);
  if (a->phases & LT_CMD_COMPUTE) {
    command_step_kernel<<<blocks, kThreads, 0, s>>>(*a, LT_CMD_COMPUTE);
    command_broadcast_kernel<<<blocks, kThreads, 0, s>>>(a->metrics, a->metric_scalars, a->N, a->gait_valid_last_air_time != nullptr);
  }
  return lt::check_launch();
}

extern "C" int lt_vel_curriculum(const LtVelCurriculumArgs* a, void* stream) {
  if (!a || a->N <= 0 || !a->ranges || !a->reset_mask || !a->episode_length_buf || !a->episode_sums_lin || !a->episode_sums_ang ||
      !a->env_reseted_lin || !a->episode_length_buf_lin || !a->episode_reward_sum_lin || !a->env_reseted_ang || !a->episode_length_buf_ang ||
      !a->episode_reward_sum_ang)
    return LT_ERR_INVALID_ARG;
  vel_curriculum_kernel<<<1, kCurThreads, 0, (cudaStream_t)stream>>>(*a);
  return lt::check_launch();
}
