// K1: fused MDP step -- terminations -> rewards (incl. the stateful adaptive symmetric gait reward) -> [auto reset]
// -> policy / critic observations with history shift and noise.
// Replaces the per-term PyTorch code of reference locotouch/mdp/rewards.py:15-604, terminations.py:10-23,
// observations.py:38-91 and the IsaacLab manager loops around it (several thousand ATen launches per env step,
// SURVEY.md 2.1) with one launch that reads exactly the IsaacLab tensors the reference terms read.
//
// Block = 256 threads = kEnvs (8) consecutive envs; the rows of 8 consecutive envs are ONE contiguous span in every
// [N, row] input tensor, so all global traffic is coalesced and the per-env arithmetic runs out of shared memory.
//   stage 0  table-driven staging: warp w copies input tensors w, w+8, ... (global -> shared); the two observation
//            history blocks (the bulk of the bytes) are fetched with 16-byte cp.async and land while stage 1 computes.
//   stage 1  warp roles, one lane per env, concurrently:
//              warp 0  gait reward (state update, task score, sync / async / stance)        rewards.py:60-392
//              warp 1  contact maxima, terminations, robot reward terms                      rewards.py:15-56, 398-466
//              warp 2  object-transport reward terms                                         rewards.py:469-604
//              warp 3  object_state_in_robot_frame for both observation groups               observations.py:38-91
//              warps 4-7  new observation values (noise, scale) of the proprioceptive terms  [IL] ObservationManager
//   stage 2  all threads: weighted accumulation in manager order, episode sums, per-term outputs, gait state write back
//            (zeroed for reset envs), and the history shift: shared -> global with 16-byte stores.
// Floating point follows the reference's fp32 expression order (file is built with -fmad=false); masks are bit-exact.
#include <cuda_pipeline.h>
#include <string.h>

#include "lt_common.cuh"

namespace {

#ifndef LT_MDP_ENVS
#define LT_MDP_ENVS 8
#endif
constexpr int kEnvs = LT_MDP_ENVS;  // envs per block (multiple of 8 so that every staged span is 16-byte aligned; <= 16)
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxObsDim = 512;
constexpr int kMaxNew = 96;
constexpr int kMaxJ = 16;
constexpr int kMaxSensorBodies = 32;
constexpr int kMaxStage = 40;
constexpr int kGaitFloats = 24;  // per env: lsa[4] lsc[4] vla[4] last_cmd[3] steps sz[4] vpc[4]

// shared-memory layout (float offsets; per-env row stride = row length), filled in on the host
struct Layout {
  int cmd, pos, linb, angb, grav, q, qd, qdd, tau, q0, qd0, lim, act, pact, force, air, con, lair, fpos, fvel;
  int quat, linw, angw, opos, oquat, olin, oang, ograv, oc_last, oc_cur, oc_air, fmax, g_lsa, g_lsc, g_vla, g_cmd, g_steps, g_sz, g_vpc, eplen, gait_out,
      esum, raw, newobs, uscr, hist, map, total;
  int D;    // observation dim per group
  int dps;  // new values per step per group
  int hist_stride;  // floats per group in the history staging area (kEnvs * D rounded up to 4)
};

// global -> shared copy descriptors for the plain [N, row] float tensors
struct StageTable {
  int n;
  const float* src[kMaxStage];
  int row[kMaxStage];
  int off[kMaxStage];
  int vec[kMaxStage];  // 1: tensor base is 16-byte aligned
};

struct Vec3 { float x, y, z; };
struct Quat { float w, x, y, z; };

__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// [IL] isaaclab.utils.math.quat_apply (sign=+1) / quat_apply_inverse (sign=-1): v + sign*w*t + xyz x t, t = 2 (xyz x v)
__device__ __noinline__ Vec3 quat_rotate(Quat q, Vec3 v, float sign) {
  const Vec3 u = {q.x, q.y, q.z};
  Vec3 t = cross(u, v);
  t = {t.x * 2.f, t.y * 2.f, t.z * 2.f};
  const Vec3 c = cross(u, t);
  const float sw = sign * q.w;
  return {v.x + sw * t.x + c.x, v.y + sw * t.y + c.y, v.z + sw * t.z + c.z};
}
__device__ __forceinline__ Vec3 rot(Quat q, Vec3 v) { return quat_rotate(q, v, 1.f); }
__device__ __forceinline__ Vec3 rot_inv(Quat q, Vec3 v) { return quat_rotate(q, v, -1.f); }
// [IL] quat_mul, 8-multiply factored form
__device__ __noinline__ Quat quat_mul(Quat a, Quat b) {
  const float ww = (a.z + a.x) * (b.x + b.y);
  const float yy = (a.w - a.y) * (b.w + b.z);
  const float zz = (a.w + a.y) * (b.w - b.z);
  const float xx = ww + yy + zz;
  const float qq = 0.5f * (xx + (a.z - a.x) * (b.x - b.y));
  return {qq - ww + (a.z - a.y) * (b.y - b.z), qq - xx + (a.x + a.w) * (b.x + b.w), qq - yy + (a.w - a.x) * (b.y + b.z),
          qq - zz + (a.z + a.y) * (b.w - b.x)};
}
// [IL] quat_inv: conj(q) / max(|q|^2, 1e-9)
__device__ __noinline__ Quat quat_inv(Quat q) {
  const float n = fmaxf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z, 1e-9f);
  return {q.w / n, -q.x / n, -q.y / n, -q.z / n};
}
__device__ __noinline__ float2 sincos_half(float angle) {
  float s, c;
  sincosf(angle * 0.5f, &s, &c);
  return make_float2(s, c);
}
__device__ __forceinline__ Quat quat_from_euler(float roll, float pitch, float yaw) {
  const float2 y = sincos_half(yaw), r = sincos_half(roll), p = sincos_half(pitch);
  const float sy = y.x, cy = y.y, sr = r.x, cr = r.y, sp = p.x, cp = p.y;
  return {cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp};
}
__device__ __noinline__ float yaw_of(Quat q) { return atan2f(2.0f * (q.w * q.z + q.x * q.y), 1.f - 2.f * (q.y * q.y + q.z * q.z)); }
__device__ __forceinline__ Quat yaw_quat(float yaw) {
  const float2 sc = sincos_half(yaw);
  return {sc.y, 0.f, 0.f, sc.x};
}
__device__ __noinline__ float exp_neg_over(float err, float sigma) { return expf(-err / sigma); }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ Vec3 ld3(const float* p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ Quat ld4(const float* p) { return {p[0], p[1], p[2], p[3]}; }
__device__ __forceinline__ Vec3 sub3(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __noinline__ float uniform_at(uint64_t seed, uint64_t offset, uint32_t n, uint32_t j) {
  const uint4 r = lt::Philox::gen(seed, offset, n, j >> 2);
  const uint32_t w = (j & 3) == 0 ? r.x : ((j & 3) == 1 ? r.y : ((j & 3) == 2 ? r.z : r.w));
  return lt::Philox::u01(w);
}

// ---------------------------------------------------------------------------------------------- gait (rewards.py:60-392)
struct GaitRegs {
  float lsa[4], lsc[4], vla[4], last_cmd[3], steps;
  bool sz[4], vpc[4];
};

// rewards.py:158-200 in statement order
__device__ __forceinline__ void gait_update(GaitRegs& g, const float a[4], const float c[4], const float la[4], Vec3 cmd, bool nz,
                                            bool any_nz, float th) {
  if (!nz) g.vla[0] = g.vla[1] = g.vla[2] = g.vla[3] = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool new_swing = g.lsa[k] < th && a[k] > th;
    if (new_swing && nz) g.sz[k] = false;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (a[k] > th && !nz) g.sz[k] = true;
  g.steps += 1.f;
  const bool chg = fabsf(cmd.x - g.last_cmd[0]) > 1.0e-3f || fabsf(cmd.y - g.last_cmd[1]) > 1.0e-3f || fabsf(cmd.z - g.last_cmd[2]) > 1.0e-3f;
  if (chg) {
    g.last_cmd[0] = cmd.x; g.last_cmd[1] = cmd.y; g.last_cmd[2] = cmd.z;
    g.steps = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { g.sz[k] = true; g.vla[k] = 0.f; }
  }
  if (any_nz) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool new_land = g.lsc[k] < th && c[k] > th;
      if (new_land && g.vpc[k] && !g.sz[k]) g.vla[k] = la[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    g.lsa[k] = a[k];
    g.lsc[k] = c[k];
    if (c[k] > th) g.vpc[k] = true;
  }
}

// scalar gait parameters passed BY VALUE to the non-inlined helpers (taking the address of a kernel parameter would make
// the compiler spill the whole parameter block to local memory)
struct GaitP {
  float judge_time_threshold, async_judge_time_threshold, air_time_gait_bound, contact_time_gait_bound;
  float tolerance_proportion, rwd_upper_bound, rwd_lower_bound, linear_scale, two_step_dt, task_performance_ratio;
  int encourage_symmetricity;
};

// rewards.py:243-346.  (a0,a1): current air times of the synced pair; (vt*, vo*): valid last air times of target / other pair
__device__ __noinline__ float gait_swing_bonus(const GaitP gp, float a0, float a1, float vt0, float vt1, float vo0, float vo1) {
  const float th = gp.judge_time_threshold;
  const float m = (a0 + a1) / 2.f;
  const bool both_air = a0 > th && a1 > th;
  const float m_t = (vt0 + vt1) / 2.f, m_o = (vo0 + vo1) / 2.f;
  const float two_dt = gp.two_step_dt;
  const bool ok_t = vt0 > th && vt1 > th && vt0 > two_dt && vt1 > two_dt;
  const bool ok_o = vo0 > th && vo1 > th && vo0 > two_dt && vo1 > two_dt;
  const bool e = both_air && (ok_t || ok_o);
  const float ref = e ? m_o : 0.f;
  const float tol = ref + gp.tolerance_proportion * ref;
  const float diff = e ? m_t - m_o : 0.f;
  const float ext = fminf(fmaxf(tol - diff, ref), tol);
  bool within = e && m <= ext;
  const bool between = e && m > ext && m <= tol;
  within = within || (e && diff < 0.f);
  const float ub = gp.rwd_upper_bound, lb = gp.rwd_lower_bound, k = gp.linear_scale;
  const float r_within = fminf(k * m, ub), r_ref = fminf(k * ref, ub), r_ext = fminf(k * ext, ub), r_tol = fminf(k * tol, ub);
  const bool lt_ = e && ext < tol;
  const float a_b = lt_ ? -r_ext / (tol - ext) : 0.f;
  const float b_b = lt_ ? -a_b * tol : 0.f;
  const float r_between = lt_ ? a_b * m + b_b : r_ext;
  const bool gt_ = e && ext > ref;
  const float a_y = gt_ ? -r_ref / (ext - ref) : 0.f;
  const float b_y = gt_ ? -a_y * tol : 0.f;
  float low = lt_ ? (diff / (gp.tolerance_proportion * ref)) * lb : r_tol;
  if (e && !ok_o) low = lb;
  low = fminf(fmaxf(low, lb), ub);
  float r_beyond = gt_ ? a_y * m + b_y : low;
  r_beyond = fmaxf(r_beyond, low);
  const float r = within ? r_within : (between ? r_between : r_beyond);
  return e ? r : 0.f;
}

// rewards.py:218-241 for the pair of gait feet (f0, f0+1)
__device__ __forceinline__ float gait_sync(const GaitP gp, const GaitRegs& g, const float a[4], const float c[4], int pair, float score) {
  const float th = gp.judge_time_threshold;
  const int f0 = pair ? 2 : 0, f1 = f0 + 1, o0 = pair ? 0 : 2;
  const bool both_air = a[f0] > th && a[f0] < gp.air_time_gait_bound && a[f1] > th && a[f1] < gp.air_time_gait_bound;
  const bool c0 = c[f0] > th && c[f0] < gp.contact_time_gait_bound;
  const bool c1 = c[f1] > th && c[f1] < gp.contact_time_gait_bound;
  const bool both_contact = c0 && c1;
  if (gp.encourage_symmetricity) {
    float bonus = gait_swing_bonus(gp, a[f0], a[f1], g.vla[f0], g.vla[f1], g.vla[o0], g.vla[o0 + 1]);
    const float scale = 1.f - gp.task_performance_ratio + gp.task_performance_ratio * score;
    if (bonus > 0.f) bonus *= scale;
    bonus += 1.f;
    return both_air ? bonus : (both_contact ? 1.f : 0.f);
  }
  return (both_air || both_contact) ? 1.f : 0.f;
}

// rewards.py:348-363
__device__ __noinline__ float gait_async(const GaitP gp, float a0, float a1, float c0t, float c1t) {
  const float th = gp.judge_time_threshold, tha = gp.async_judge_time_threshold;
  const bool both = c0t > th && c0t <= tha && c1t > th && c1t <= tha;
  const bool air0 = a0 > th && a0 < gp.air_time_gait_bound, air1 = a1 > th && a1 < gp.air_time_gait_bound;
  const bool con0 = c0t > th && c0t < gp.contact_time_gait_bound, con1 = c1t > th && c1t < gp.contact_time_gait_bound;
  return (both || (air0 && con1) || (con0 && air1)) ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------------------- kernels
__global__ void any_nonzero_cmd_kernel(const float* __restrict__ cmd, int N, int* flag_ws, int step) {
  bool nz = false;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float x = cmd[3 * n], y = cmd[3 * n + 1], z = cmd[3 * n + 2];
    nz = nz || (sqrtf(x * x + y * y + z * z) > 0.f);
  }
  const int any = __syncthreads_or(nz ? 1 : 0);
  if (threadIdx.x == 0) flag_ws[step & 1] = any ? step : -1;
}

__global__ void __launch_bounds__(kThreads, kEnvs <= 8 ? 4 : 2) mdp_step_kernel(const LtMdpArgs A, const Layout L, const StageTable ST) {
  extern __shared__ __align__(16) float sm[];
  __shared__ unsigned char s_done[kEnvs], s_fill[kEnvs];
  __shared__ int s_step, s_any_nz;
  __shared__ signed char s_slot[LT_RK_COUNT];  // reward kind -> index in the term table (-1: absent or zero weight)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e0 = blockIdx.x * kEnvs;
  const int nvalid = min(kEnvs, A.N - e0);
  const bool do_rew = A.phases & LT_PHASE_REWARDS, do_obs = A.phases & LT_PHASE_OBS;
  const bool has_obj = A.obj_root_pos_w != nullptr;
  const int J = A.J, S = A.num_sensor_bodies, H = A.force_history;
  const int D = L.D, dps = L.dps, T = A.num_reward_terms;

  // ------------------------------------------------------------------------------------------------ stage 0: loads
  // (b) plain [N, row] tensors, one descriptor per warp and round; full blocks move 16 bytes per lane (8 envs x row floats is
  //     a multiple of 16 bytes and the span starts 16-byte aligned whenever the tensor itself does)
#pragma unroll 1
  for (int t = warp; t < ST.n; t += kWarps) {
    const int row = ST.row[t];
    const float* src = ST.src[t] + (size_t)e0 * row;
    float* dst = sm + ST.off[t];
    const int count = nvalid * row;
    if (ST.vec[t] && nvalid == kEnvs) {
      for (int i = lane; i < (count >> 2); i += 32) __pipeline_memcpy_async(dst + 4 * i, src + 4 * i, 16);
    } else {
      for (int i = lane; i < count; i += 32) __pipeline_memcpy_async(dst + i, src + i, 4);
    }
  }
  // (c) gathers with arbitrary body ids / non-float types: ONE pass, every thread issues at most a few independent
  //     4-byte cp.async (no load sits inside a loop-carried chain, so the block pays a single DRAM round trip)
  if (do_rew) {
    if (tid < nvalid * 4) {  // feet rows of the contact timers
      const int e = tid >> 2, k = tid & 3;
      const size_t r = (size_t)(e0 + e) * S + A.gait.feet_ids[k];
      __pipeline_memcpy_async(sm + L.air + tid, A.current_air_time + r, 4);
      __pipeline_memcpy_async(sm + L.con + tid, A.current_contact_time + r, 4);
      __pipeline_memcpy_async(sm + L.lair + tid, A.last_air_time + r, 4);
    }
    if (tid >= 32 && tid < 32 + nvalid * 12) {  // feet rows of body_pos_w / body_lin_vel_w
      const int i = tid - 32;
      const int e = i / 12, k = (i % 12) / 3, c = i % 3;
      const size_t r = ((size_t)(e0 + e) * A.num_bodies + A.feet_body_ids[k]) * 3 + c;
      __pipeline_memcpy_async(sm + L.fpos + i, A.body_pos_w + r, 4);
      __pipeline_memcpy_async(sm + L.fvel + i, A.body_lin_vel_w + r, 4);
    }
    if (A.episode_sums && tid >= 128 && tid < 128 + T) {  // [terms][N]: kEnvs contiguous floats per term
      const int t = tid - 128;
      const float* src = A.episode_sums + (size_t)t * A.N + e0;
      for (int e = 0; e < nvalid; ++e) __pipeline_memcpy_async(sm + L.esum + t * kEnvs + e, src + e, 4);
    }
    if (tid >= 192 && tid < 192 + nvalid * 4) {  // bool gait state -> float
      const int i = tid - 192;
      sm[L.g_sz + i] = (float)A.gait_state.swinging_in_zero_cmd[(size_t)e0 * 4 + i];
      sm[L.g_vpc + i] = (float)A.gait_state.valid_previous_contact[(size_t)e0 * 4 + i];
    }
    if (tid >= 224 && tid < 224 + nvalid) {
      reinterpret_cast<long long*>(sm + L.eplen)[tid - 224] = A.episode_length_buf[e0 + tid - 224];
    }
  }
  if (tid == 255) {  // env-step index and the cross-env any(non_zero_cmd) flag (one dependent pair of loads, one thread)
    const int st = (int)(A.offset + (A.offset_base ? (uint64_t)*A.offset_base : 0ull));
    s_step = st;
    int any_nz = 1;
    if (A.any_nonzero_cmd_override >= 0) any_nz = A.any_nonzero_cmd_override != 0;
    else if (A.any_flag_ws) any_nz = A.any_flag_ws[st & 1] == st;
    s_any_nz = any_nz;
  }
  __pipeline_commit();
  // (a) observation history blocks: the block's kEnvs rows of each group are one 16-byte aligned contiguous span
  if (do_obs) {
    const int total = nvalid * D;
#pragma unroll 1
    for (int grp = 0; grp < 2; ++grp) {
      const float* in = grp ? A.critic_obs_in : A.policy_obs_in;
      if (!in) continue;
      const float* src = in + (size_t)e0 * D;
      float* dst = sm + L.hist + grp * L.hist_stride;
      if ((((uintptr_t)src) & 15) == 0) {
        const int n4 = total >> 2;
        for (int i = tid; i < n4; i += kThreads) __pipeline_memcpy_async(dst + 4 * i, src + 4 * i, 16);
        for (int i = (n4 << 2) + tid; i < total; i += kThreads) __pipeline_memcpy_async(dst + i, src + i, 4);
      } else {
        for (int i = tid; i < total; i += kThreads) __pipeline_memcpy_async(dst + i, src + i, 4);
      }
    }
    __pipeline_commit();
  }
  // Observation layout tables (built once per block):
  //   s_map[k]  for column k of the flattened [term][history][dim] row: low 16 bits = index j of the per-step value that
  //             feeds the column (used when the history is (re)filled), high 16 bits = k + d, the column one history
  //             slot later of the same term (the shift source), or 0xffff when k is the newest slot
  //   s_jinfo[j] for per-step value j: term index | component << 8
  int* s_map = reinterpret_cast<int*>(sm + L.map);
  int* s_jinfo = s_map + ((D + 3) & ~3);
  if (do_obs && warp >= 5) {
    for (int k = tid - 5 * 32; k < D + dps; k += 3 * 32) {
      int col = 0, jbase = 0;
      if (k < D) {
        int entry = 0;
        for (int t = 0; t < A.num_obs_terms; ++t) {
          const int d = A.obs_terms[t].dim, span = d * A.history_length;
          if (k < col + span) {
            const int h = (k - col) / d, i = (k - col) % d;
            entry = (jbase + i) | ((h == A.history_length - 1 ? 0xffff : k + d) << 16);
            break;
          }
          col += span;
          jbase += d;
        }
        s_map[k] = entry;
      } else {
        const int j = k - D;
        int t = 0;
        while (j >= jbase + A.obs_terms[t].dim) { jbase += A.obs_terms[t].dim; ++t; }
        s_jinfo[j] = t | ((j - jbase) << 8);
      }
    }
  }
  if (do_rew && tid >= 64 && tid < 64 + LT_RK_COUNT) {
    const int kind = tid - 64;
    int slot = -1;
#pragma unroll 1
    for (int i = 0; i < T; ++i)
      if (A.reward_terms[i].kind == kind && A.reward_terms[i].weight != 0.f) slot = i;
    s_slot[kind] = (signed char)slot;
  }
  __pipeline_wait_prior(do_obs ? 1 : 0);  // the state tensors have landed; the history blocks may still be in flight
  if (tid < kEnvs) {
    s_fill[tid] = (do_obs && A.obs_fill && tid < nvalid) ? A.obs_fill[e0 + tid] : 0;
    s_done[tid] = 0;
  }
  __syncthreads();
  const int step = s_step;
  const uint64_t rng_offset = (uint64_t)(int64_t)step;

  // ------------------------------------------------------------------------------------------------ stage 1: roles
  float* s_raw = sm + L.raw;  // [T][kEnvs]
  if (warp == 0) {
    // ---- gait reward
    if (do_rew && lane < nvalid) {
      const int e = lane;
      const int gi = s_slot[LT_RK_GAIT];
      if (gi >= 0) {
        GaitP gp;
        gp.judge_time_threshold = A.gait.judge_time_threshold;
        gp.async_judge_time_threshold = A.gait.async_judge_time_threshold;
        gp.air_time_gait_bound = A.gait.air_time_gait_bound;
        gp.contact_time_gait_bound = A.gait.contact_time_gait_bound;
        gp.tolerance_proportion = A.gait.tolerance_proportion;
        gp.rwd_upper_bound = A.gait.rwd_upper_bound;
        gp.rwd_lower_bound = A.gait.rwd_lower_bound;
        gp.linear_scale = A.gait.linear_scale;
        gp.two_step_dt = A.gait.two_step_dt;
        gp.task_performance_ratio = A.gait.task_performance_ratio;
        gp.encourage_symmetricity = A.gait.encourage_symmetricity;
        const float th = gp.judge_time_threshold;
        const Vec3 cmd = ld3(sm + L.cmd + e * 3), vb = ld3(sm + L.linb + e * 3);
        const float wz = sm[L.angb + e * 3 + 2];
        const bool nz = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z) > 0.f;
        GaitRegs g;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          g.lsa[k] = sm[L.g_lsa + e * 4 + k]; g.lsc[k] = sm[L.g_lsc + e * 4 + k]; g.vla[k] = sm[L.g_vla + e * 4 + k];
          g.sz[k] = sm[L.g_sz + e * 4 + k] != 0.f; g.vpc[k] = sm[L.g_vpc + e * 4 + k] != 0.f;
        }
        g.last_cmd[0] = sm[L.g_cmd + e * 3]; g.last_cmd[1] = sm[L.g_cmd + e * 3 + 1]; g.last_cmd[2] = sm[L.g_cmd + e * 3 + 2];
        g.steps = sm[L.g_steps + e];
        float a[4], c[4], la[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = sm[L.air + e * 4 + k]; c[k] = sm[L.con + e * 4 + k]; la[k] = sm[L.lair + e * 4 + k]; }
        const bool any_nz = s_any_nz != 0;
        gait_update(g, a, c, la, cmd, nz, any_nz, th);
        float score = 0.f;  // rewards.py:202-216 / 372-392
        if (gp.encourage_symmetricity) {
          const float dx = cmd.x - vb.x, dy = cmd.y - vb.y;
          const float e_lin = nz ? sqrtf(dx * dx + dy * dy) : 0.f;
          const float e_ang = nz ? fabsf(cmd.z - wz) : 0.f;
          score = (exp_neg_over(e_lin, A.gait.vel_tracking_exp_sigma) + exp_neg_over(e_ang, A.gait.vel_tracking_exp_sigma)) / 2.f;
          if (A.gait.with_object) {
            const Vec3 rel_w = sub3(ld3(sm + L.opos + e * 3), ld3(sm + L.pos + e * 3));
            const Vec3 r = rot_inv(yaw_quat(yaw_of(ld4(sm + L.quat + e * 4))), rel_w);
            const float bx = clampf(1.f - fabsf(r.x) / A.gait.obj_x_max, 0.f, 1.f);
            const float by = clampf(1.f - fabsf(r.y) / A.gait.obj_y_max, 0.f, 1.f);
            score = clampf((score * 2.f + (bx + by) / 2.f) / 3.f, 0.f, 1.f);
          }
        }
        const float sync = (gait_sync(gp, g, a, c, 0, score) + gait_sync(gp, g, a, c, 1, score)) / 2.f;
        const float asyn = (gait_async(gp, a[0], a[2], c[0], c[2]) + gait_async(gp, a[1], a[3], c[1], c[3]) +
                            gait_async(gp, a[0], a[3], c[0], c[3]) + gait_async(gp, a[2], a[1], c[2], c[1])) / 4.f;
        const float stepping = (sync + asyn) / 2.f;
        const float stance = ((c[0] > th && c[1] > th && c[2] > th && c[3] > th) ? 1.f : 0.f) * A.gait.stance_rwd_scale;
        s_raw[gi * kEnvs + e] = nz ? stepping : stance;
        float* go = sm + L.gait_out + e * kGaitFloats;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          go[k] = g.lsa[k]; go[4 + k] = g.lsc[k]; go[8 + k] = g.vla[k];
          go[16 + k] = g.sz[k] ? 1.f : 0.f; go[20 + k] = g.vpc[k] ? 1.f : 0.f;
        }
        go[12] = g.last_cmd[0]; go[13] = g.last_cmd[1]; go[14] = g.last_cmd[2]; go[15] = g.steps;
      } else {
        // no gait term: carry the state through unchanged
        float* go = sm + L.gait_out + e * kGaitFloats;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          go[k] = sm[L.g_lsa + e * 4 + k]; go[4 + k] = sm[L.g_lsc + e * 4 + k]; go[8 + k] = sm[L.g_vla + e * 4 + k];
          go[16 + k] = sm[L.g_sz + e * 4 + k]; go[20 + k] = sm[L.g_vpc + e * 4 + k];
        }
        go[12] = sm[L.g_cmd + e * 3]; go[13] = sm[L.g_cmd + e * 3 + 1]; go[14] = sm[L.g_cmd + e * 3 + 2]; go[15] = sm[L.g_steps + e];
      }
    }
  } else if (warp == 1) {
    // ---- contact maxima, terminations, robot terms
    if (do_rew) {
      for (int i = lane; i < nvalid * S; i += 32) {
        const int e = i / S, b = i % S;
        const float* f = sm + L.force + e * (H * S * 3);
        float m = 0.f;
#pragma unroll 1
        for (int h = 0; h < H; ++h) {
          const float* p = f + (h * S + b) * 3;
          m = fmaxf(m, sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));  // torch.max(norm(F, dim=-1), dim=1)
        }
        sm[L.fmax + i] = m;
      }
      __syncwarp();
      if (lane < nvalid) {
        const int e = lane, n = e0 + e;
        bool terminated = false, timed_out = false;
#pragma unroll 1
        for (int t = 0; t < A.num_termination_terms; ++t) {
          const LtTerminationTerm& tt = A.termination_terms[t];
          bool m = false;
          switch (tt.kind) {
            case LT_TK_TIME_OUT: m = reinterpret_cast<const long long*>(sm + L.eplen)[e] >= A.max_episode_length; break;
            case LT_TK_BAD_ORIENTATION: m = fabsf(acosf(-sm[L.grav + e * 3 + 2])) > tt.p[0]; break;
            case LT_TK_ROOT_HEIGHT: m = sm[L.pos + e * 3 + 2] < tt.p[0]; break;
            case LT_TK_ILLEGAL_CONTACT:
#pragma unroll 1
              for (int k = 0; k < tt.num_ids; ++k) m = m || sm[L.fmax + e * S + tt.body_ids[k]] > tt.p[0];
              break;
            case LT_TK_OBJECT_BELOW_ROBOT: m = sm[L.opos + e * 3 + 2] < sm[L.pos + e * 3 + 2]; break;
            case LT_TK_BAD_ROLL: m = fabsf(asinf(sm[L.ograv + e * 3 + 1])) > tt.p[0]; break;
          }
          if (A.term_masks) A.term_masks[(size_t)t * A.N + n] = m;
          if (tt.time_out) timed_out = timed_out || m; else terminated = terminated || m;
        }
        const bool done = terminated || timed_out;
        s_done[e] = done;
        A.terminated[n] = terminated;
        A.time_outs[n] = timed_out;
        A.dones[n] = done;
        if (A.auto_reset && done) {  // history of a reset env is refilled by its next observation
          if (do_obs) s_fill[e] = 1;
          else if (A.obs_fill) A.obs_fill[n] = 1;
        }
        // robot reward terms, straight-line: every value is computed, stored only when the task lists the term
        auto put = [&](int kind, float v) {
          const int i = s_slot[kind];
          if (i >= 0) s_raw[i * kEnvs + e] = v;
        };
        auto par = [&](int kind, int k) -> float {
          const int i = s_slot[kind];
          return A.reward_terms[i < 0 ? 0 : i].p[k];
        };
        const Vec3 cmd = ld3(sm + L.cmd + e * 3), vb = ld3(sm + L.linb + e * 3), wb = ld3(sm + L.angb + e * 3);
        const Vec3 grav = ld3(sm + L.grav + e * 3);
        const float cmd_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z);
        const float* fmx = sm + L.fmax + e * S;
        const float* fpos = sm + L.fpos + e * 12;
        const float* fvel = sm + L.fvel + e * 12;
        put(LT_RK_ALIVE, terminated ? 0.f : 1.f);
        {
          const float dx = cmd.x - vb.x, dy = cmd.y - vb.y;
          put(LT_RK_TRACK_LIN_VEL_XY, exp_neg_over(sqrtf(dx * dx + dy * dy), par(LT_RK_TRACK_LIN_VEL_XY, 0)));
          put(LT_RK_TRACK_ANG_VEL_Z, exp_neg_over(fabsf(cmd.z - wb.z), par(LT_RK_TRACK_ANG_VEL_Z, 0)));
        }
        {
          const float slip_thr = par(LT_RK_FOOT_SLIP, 0), drag_h = par(LT_RK_FOOT_DRAG, 0), drag_v = par(LT_RK_FOOT_DRAG, 1);
          float slip = 0.f, drag = 0.f;
#pragma unroll 1
          for (int k = 0; k < 4; ++k) {
            const float sp = sqrtf(fvel[3 * k] * fvel[3 * k] + fvel[3 * k + 1] * fvel[3 * k + 1]);
            slip += (fmx[A.feet_sensor_ids[k]] > slip_thr ? 1.f : 0.f) * sp;
            drag += (fpos[3 * k + 2] <= drag_h && sp > drag_v) ? 1.f : 0.f;
          }
          put(LT_RK_FOOT_SLIP, slip);
          put(LT_RK_FOOT_DRAG, drag);
        }
        {
          const float d = sm[L.pos + e * 3 + 2] - par(LT_RK_BASE_HEIGHT, 0);
          put(LT_RK_BASE_HEIGHT, d * d);
          put(LT_RK_BASE_Z_VEL, vb.z * vb.z);
          put(LT_RK_BASE_RP_ANGLE, grav.x * grav.x + grav.y * grav.y);
          put(LT_RK_BASE_RP_VEL, fabsf(wb.x) + fabsf(wb.y));
        }
        {
          const float* qq = sm + L.q + e * J; const float* q0 = sm + L.q0 + e * J; const float* lim = sm + L.lim + e * 2 * J;
          const float* qd = sm + L.qd + e * J; const float* qdd = sm + L.qdd + e * J; const float* tau = sm + L.tau + e * J;
          const float* ac = sm + L.act + e * J; const float* pa = sm + L.pact + e * J;
          float s_lim = 0.f, s_dev = 0.f, s_acc = 0.f, s_vel = 0.f, s_tau = 0.f, s_rate = 0.f;
#pragma unroll 2
          for (int j = 0; j < J; ++j) {
            s_lim += -fminf(qq[j] - lim[2 * j], 0.f) + fmaxf(qq[j] - lim[2 * j + 1], 0.f);
            const float d = qq[j] - q0[j];
            s_dev += d * d;
            s_acc += qdd[j] * qdd[j];
            s_vel += qd[j] * qd[j];
            s_tau += tau[j] * tau[j];
            const float da = ac[j] - pa[j];
            s_rate += da * da;
          }
          put(LT_RK_JOINT_POS_LIMIT, s_lim);
          const float dev = sqrtf(s_dev), bv = sqrtf(vb.x * vb.x + vb.y * vb.y);
          put(LT_RK_JOINT_POS, (cmd_norm > 0.f || bv > par(LT_RK_JOINT_POS, 1)) ? dev : par(LT_RK_JOINT_POS, 0) * dev);
          put(LT_RK_JOINT_ACC, sqrtf(s_acc));
          put(LT_RK_JOINT_VEL, sqrtf(s_vel));
          put(LT_RK_JOINT_TORQUE, sqrtf(s_tau));
          put(LT_RK_ACTION_RATE, s_rate);
        }
        {
          const float thr = par(LT_RK_THIGH_CALF_COLLISION, 0);
          float c = 0.f;
#pragma unroll 1
          for (int k = 0; k < A.num_thigh_calf; ++k) c += fmx[A.thigh_calf_sensor_ids[k]] > thr ? 1.f : 0.f;
          put(LT_RK_THIGH_CALF_COLLISION, c);
        }
      }
    }
  } else if (warp == 2) {
    // ---- object-transport terms
    if (do_rew && has_obj && lane < nvalid) {
      const int e = lane;
      const Vec3 cmd = ld3(sm + L.cmd + e * 3);
      const float cmd_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z);
      const Quat q = ld4(sm + L.quat + e * 4);
      const Vec3 rel_pos_w = sub3(ld3(sm + L.opos + e * 3), ld3(sm + L.pos + e * 3));
      const Vec3 rel_pos = rot_inv(q, rel_pos_w);
      const Vec3 rel_vel = rot_inv(q, sub3(ld3(sm + L.olin + e * 3), ld3(sm + L.linw + e * 3)));
      const Vec3 rel_ang = rot_inv(q, sub3(ld3(sm + L.oang + e * 3), ld3(sm + L.angw + e * 3)));
      const Vec3 g_obj = rot_inv(q, rot(ld4(sm + L.oquat + e * 4), ld3(sm + L.ograv + e * 3)));
      auto put = [&](int kind, float v) {
        const int i = s_slot[kind];
        if (i >= 0) s_raw[i * kEnvs + e] = v;
      };
      auto par = [&](int kind, int k) -> float {
        const int i = s_slot[kind];
        return A.reward_terms[i < 0 ? 0 : i].p[k];
      };
      const float moving = cmd_norm > 0.f ? 1.f : 0.f;
      {
        float v = sqrtf(rel_pos_w.x * rel_pos_w.x + rel_pos_w.y * rel_pos_w.y);
        if (par(LT_RK_OBJ_XY_POS, 0) != 0.f) v *= moving;
        put(LT_RK_OBJ_XY_POS, v);
      }
      put(LT_RK_OBJ_XY_VEL, rel_vel.x * rel_vel.x + rel_vel.y * rel_vel.y);
      put(LT_RK_OBJ_LOSE_CONTACT, (sm[L.oc_last + e] > 0.f && sm[L.oc_air + e] > 0.f) ? 1.f : 0.f);
      put(LT_RK_OBJ_Z_VEL, rel_vel.z * rel_vel.z);
      put(LT_RK_OBJ_RP_ANGLE, g_obj.x * g_obj.x + g_obj.y * g_obj.y);
      put(LT_RK_OBJ_RP_VEL, fabsf(rel_ang.x) + fabsf(rel_ang.y));
      put(LT_RK_OBJ_ROLL_ANGLE, g_obj.y * g_obj.y);
      put(LT_RK_OBJ_ROLL_VEL, rel_ang.x * rel_ang.x);
      if (s_slot[LT_RK_OBJ_YAW] >= 0) {  // rewards.py:545-567
        const Quat qr = yaw_quat(yaw_of(q)), qo = yaw_quat(yaw_of(ld4(sm + L.oquat + e * 4)));
        float d = yaw_of(quat_mul(quat_inv(qr), qo));
        const float pi = 3.14159274101257324f;  // float32(torch.pi)
        if (d > pi) d -= 2.f * pi;
        if (d > 0.5f * pi) d -= pi;
        if (d <= -0.5f * pi) d += pi;
        float v = d * d;
        if (par(LT_RK_OBJ_YAW, 0) != 0.f) v *= moving;
        put(LT_RK_OBJ_YAW, v);
      }
      if (s_slot[LT_RK_OBJ_DANGER] >= 0) {  // rewards.py:569-594
        bool bad = fabsf(rel_pos.x) > par(LT_RK_OBJ_DANGER, 0);
        bad = bad || fabsf(rel_pos.y) > par(LT_RK_OBJ_DANGER, 1);
        bad = bad || rel_pos.z < par(LT_RK_OBJ_DANGER, 2);
        const float rp = par(LT_RK_OBJ_DANGER, 3), vmax = par(LT_RK_OBJ_DANGER, 4);
        if (rp >= 0.f) bad = bad || fabsf(acosf(-sm[L.ograv + e * 3 + 2])) > rp * 3.14159265358979323846f / 180.f;
        if (vmax >= 0.f) bad = bad || sqrtf(rel_vel.x * rel_vel.x + rel_vel.y * rel_vel.y) > vmax;
        put(LT_RK_OBJ_DANGER, bad ? 1.f : 0.f);
      }
    }
  } else if (warp == 3) {
    // ---- object_state_in_robot_frame (observations.py:38-91): lane = 2*env + group
    if (do_obs && has_obj && lane < nvalid * 2) {
      int jb = 0, tobj = -1;
      for (int t = 0; t < A.num_obs_terms; ++t) {
        if (A.obs_terms[t].kind == LT_OK_OBJECT_STATE) { tobj = t; break; }
        jb += A.obs_terms[t].dim;
      }
      if (tobj >= 0) {
        const int e = lane >> 1, grp = lane & 1, n = e0 + e;
        float* nv = sm + L.newobs;
        const Quat q = ld4(sm + L.quat + e * 4);
        const Vec3 p = rot_inv(q, sub3(ld3(sm + L.opos + e * 3), ld3(sm + L.pos + e * 3)));
        const Vec3 v = rot_inv(q, sub3(ld3(sm + L.olin + e * 3), ld3(sm + L.linw + e * 3)));
        const Quat qr = quat_mul(quat_inv(q), ld4(sm + L.oquat + e * 4));
        const Vec3 w = rot_inv(q, sub3(ld3(sm + L.oang + e * 3), ld3(sm + L.angw + e * 3)));
        float st[13] = {p.x, p.y, p.z, v.x, v.y, v.z, qr.w, qr.x, qr.y, qr.z, w.x, w.y, w.z};
        float cst[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) cst[k] = A.os_non_contact[k];
        if (grp == 0) {  // policy group: add_uniform_noise=True
          float* u13 = sm + L.uscr + e * 16;  // 13 uniforms of this env (one policy-group lane per env)
          if (A.u_obs) {
#pragma unroll 1
            for (int k = 0; k < 13; ++k) u13[k] = __ldcs(A.u_obs + (size_t)n * dps + jb + k);
          } else {  // the same (env, quad) counter stream the proprioceptive values use: value j takes word j & 3 of quad j >> 2
#pragma unroll 1
            for (int qd = jb >> 2; qd <= (jb + 12) >> 2; ++qd) {
              const uint4 r = lt::Philox::gen(A.seed, rng_offset, (uint32_t)n, (uint32_t)qd);
              const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                const int k = 4 * qd + c4 - jb;
                if (k >= 0 && k < 13) u13[k] = lt::Philox::u01(rw[c4]);
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 13; ++k) {
            const float add = u13[k] * (A.os_n_max[k] - A.os_n_min[k]) + A.os_n_min[k];  // observations.py:77
            st[k] = st[k] + add;
            cst[k] = cst[k] + add;  // observations.py:82 (same draw)
          }
          float de[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float u = A.u_obj_euler ? A.u_obj_euler[n * 3 + k] : uniform_at(A.seed, rng_offset, (uint32_t)n, 0x1000u + k);
            de[k] = u * (A.os_euler_max[k] - A.os_euler_min[k]) + A.os_euler_min[k];
          }
          const Quat nq = quat_from_euler(de[0], de[1], de[2]);  // observations.py:78-79
          const Quat a = quat_mul({st[6], st[7], st[8], st[9]}, nq);
          st[6] = a.w; st[7] = a.x; st[8] = a.y; st[9] = a.z;
          const Quat b = quat_mul({cst[6], cst[7], cst[8], cst[9]}, nq);
          cst[6] = b.w; cst[7] = b.x; cst[8] = b.y; cst[9] = b.z;
        }
        const bool never = sm[L.oc_last + e] < A.os_last_contact_thr && sm[L.oc_cur + e] < A.os_current_contact_thr;
        const float term_scale = A.obs_terms[tobj].scale;
#pragma unroll
        for (int k = 0; k < 13; ++k) nv[(grp * kEnvs + e) * dps + jb + k] = ((never ? cst[k] : st[k]) * A.os_scale[k]) * term_scale;
      }
    }
  } else {
    // ---- warps 4..7: new observation values of the proprioceptive terms; one thread per (env, 4 consecutive values)
    if (do_obs) {
      float* nv = sm + L.newobs;  // [2 groups][kEnvs][dps]
      const int quads = (dps + 3) >> 2;
      for (int i = tid - 4 * 32; i < nvalid * quads; i += 4 * 32) {
        const int e = i / quads, qd = i - e * quads, n = e0 + e;
        uint4 rnd = make_uint4(0, 0, 0, 0);
        if (!A.u_obs) rnd = lt::Philox::gen(A.seed, rng_offset, (uint32_t)n, (uint32_t)qd);
        const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int j = 4 * qd + c4;
          if (j >= dps) break;
          const int info = s_jinfo[j];
          const LtObsTerm& ot = A.obs_terms[info & 0xff];
          if (ot.kind == LT_OK_OBJECT_STATE) continue;
          const int c = info >> 8;
          float raw;
          switch (ot.kind) {
            case LT_OK_COMMAND: raw = sm[L.cmd + e * 3 + c]; break;
            case LT_OK_BASE_ANG_VEL: raw = sm[L.angb + e * 3 + c]; break;
            case LT_OK_PROJECTED_GRAVITY: raw = sm[L.grav + e * 3 + c]; break;
            case LT_OK_JOINT_POS_REL: raw = sm[L.q + e * J + c] - sm[L.q0 + e * J + c]; break;
            case LT_OK_JOINT_VEL_REL: raw = sm[L.qd + e * J + c] - sm[L.qd0 + e * J + c]; break;
            default: raw = sm[L.act + e * J + c]; break;  // LT_OK_LAST_ACTION
          }
          float noisy = raw;
          if (ot.noisy) {
            const float u = A.u_obs ? __ldcs(A.u_obs + (size_t)n * dps + j) : lt::Philox::u01(rw[c4]);
            noisy = (raw + u * (ot.n_max - ot.n_min)) + ot.n_min;  // [IL] data + rand*(max-min) + min
          }
          nv[(0 * kEnvs + e) * dps + j] = noisy * ot.scale;
          nv[(1 * kEnvs + e) * dps + j] = raw * ot.scale;
        }
      }
      // publish any(non_zero_cmd) for the reward pass of the NEXT step (it sees the same command tensor)
      if (A.any_flag_ws && warp == 4 && lane < nvalid) {
        const float* c = sm + L.cmd + lane * 3;
        if (sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) > 0.f) A.any_flag_ws[(step + 1) & 1] = step + 1;
      }
    }
  }
  if (do_obs) __pipeline_wait_prior(0);
  __syncthreads();

  // ---------------------------------------------------------------------------- stage 2a: reward accumulation + outputs
  if (do_rew) {
    const float dt = A.step_dt;
    // [IL] RewardManager.compute: value = f * weight * dt ; episode_sums += value ; step_reward = value / dt
    for (int i = tid; i < T * kEnvs; i += kThreads) {
      const int t = i / kEnvs, e = i % kEnvs, n = e0 + e;
      if (e >= nvalid) continue;
      const LtRewardTerm& rt = A.reward_terms[t];
      if (rt.weight == 0.f) {  // skipped terms: step_reward column is zero, sums untouched
        if (A.step_reward) A.step_reward[(size_t)n * T + t] = 0.f;
        continue;
      }
      const float raw = s_raw[t * kEnvs + e];
      const float value = (raw * rt.weight) * dt;
      if (A.term_raw) A.term_raw[(size_t)t * A.N + n] = raw;
      if (A.step_reward) A.step_reward[(size_t)n * T + t] = value / dt;
      if (A.episode_sums) {
        const float total = sm[L.esum + t * kEnvs + e] + value;
        const bool rst = A.auto_reset && s_done[e];
        if (rst && A.episode_log_sums) atomicAdd(A.episode_log_sums + t, total);
        A.episode_sums[(size_t)t * A.N + n] = rst ? 0.f : total;
      }
    }
    if (warp == kWarps - 1 && lane < nvalid) {  // reward_buf: sequential fp32 sum in term order
      const int e = lane;
      float reward = 0.f;
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        const LtRewardTerm& rt = A.reward_terms[t];
        if (rt.weight != 0.f) reward = reward + (s_raw[t * kEnvs + e] * rt.weight) * dt;
      }
      A.reward[e0 + e] = reward;
      if (A.auto_reset && s_done[e] && A.episode_log_sums) atomicAdd(A.episode_log_sums + T, 1.0f);
    }
    // gait state write back; zeroed when the env is being reset (the manager's reset(env_ids) follows, rewards.py:107-114)
    {
      const LtGaitState& G = A.gait_state;
      const float* go = sm + L.gait_out;
      for (int i = tid; i < nvalid * kGaitFloats; i += kThreads) {
        const int e = i / kGaitFloats, k = i % kGaitFloats, n = e0 + e;
        const float v = (A.auto_reset && s_done[e]) ? 0.f : go[i];
        if (k < 4) G.last_step_current_air_time[n * 4 + k] = v;
        else if (k < 8) G.last_step_current_contact_time[n * 4 + k - 4] = v;
        else if (k < 12) G.valid_last_air_time[n * 4 + k - 8] = v;
        else if (k < 15) G.last_velocity_cmd[n * 3 + k - 12] = v;
        else if (k == 15) G.step_from_changing_cmd[n] = v;
        else if (k < 20) G.swinging_in_zero_cmd[n * 4 + k - 16] = v != 0.f;
        else G.valid_previous_contact[n * 4 + k - 20] = v != 0.f;
      }
    }
  }

  // --------------------------------------------------------------------------------- stage 2b: observation history shift
  // warp w owns env row w: rows are 8-byte aligned (D is even), so every lane moves float2 pairs
  if (do_obs) {
    const float* nv = sm + L.newobs;
    for (int e = warp; e < nvalid; e += kWarps) {
      const bool fill = s_fill[e] != 0;
#pragma unroll 1
      for (int grp = 0; grp < 2; ++grp) {
        float* out = grp ? A.critic_obs_out : A.policy_obs_out;
        if (!out) continue;
        const bool refill = fill || (grp ? A.critic_obs_in : A.policy_obs_in) == nullptr;
        const float* hist = sm + L.hist + grp * L.hist_stride + e * D;
        const float* nve = nv + (grp * kEnvs + e) * dps;
        float* dst = out + (size_t)(e0 + e) * D;
        if ((D & 1) == 0 && (((uintptr_t)dst) & 7) == 0) {
          for (int p = lane; p < (D >> 1); p += 32) {
            const uint2 m = *reinterpret_cast<const uint2*>(s_map + 2 * p);
            float2 v;
            v.x = (refill || (m.x >> 16) == 0xffff) ? nve[m.x & 0xffff] : hist[m.x >> 16];
            v.y = (refill || (m.y >> 16) == 0xffff) ? nve[m.y & 0xffff] : hist[m.y >> 16];
            __stcs(reinterpret_cast<float2*>(dst) + p, v);
          }
        } else {
          for (int k = lane; k < D; k += 32) {
            const unsigned m = (unsigned)s_map[k];
            __stcs(dst + k, (refill || (m >> 16) == 0xffff) ? nve[m & 0xffff] : hist[m >> 16]);
          }
        }
      }
      if (A.obs_fill && lane == 0) A.obs_fill[e0 + e] = 0;
    }
  }
}

__global__ void mdp_reset_kernel(const LtGaitState G, float* episode_sums, int num_terms, const uint8_t* mask, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N || !mask[n]) return;
  for (int k = 0; k < 4; ++k) {
    G.last_step_current_air_time[n * 4 + k] = 0.f;
    G.last_step_current_contact_time[n * 4 + k] = 0.f;
    G.valid_last_air_time[n * 4 + k] = 0.f;
    G.swinging_in_zero_cmd[n * 4 + k] = 0;
    G.valid_previous_contact[n * 4 + k] = 0;
  }
  for (int k = 0; k < 3; ++k) G.last_velocity_cmd[n * 3 + k] = 0.f;
  G.step_from_changing_cmd[n] = 0.f;
  if (episode_sums)
    for (int i = 0; i < num_terms; ++i) episode_sums[(size_t)i * N + n] = 0.f;
}

}  // namespace

extern "C" int lt_mdp_step(const LtMdpArgs* a, void* stream) {
  if (!a || a->N <= 0 || !(a->phases & (LT_PHASE_REWARDS | LT_PHASE_OBS))) return LT_ERR_INVALID_ARG;
  const bool do_rew = a->phases & LT_PHASE_REWARDS, do_obs = a->phases & LT_PHASE_OBS;
  const bool has_obj = a->obj_root_pos_w != nullptr;
  if (a->J <= 0 || a->J > kMaxJ) return LT_ERR_INVALID_ARG;
  if (!a->command || !a->root_pos_w || !a->root_ang_vel_b || !a->projected_gravity_b || !a->joint_pos || !a->joint_vel ||
      !a->default_joint_pos || !a->raw_actions)
    return LT_ERR_INVALID_ARG;
  if (do_rew) {
    if (!a->root_lin_vel_b || !a->joint_acc || !a->applied_torque || !a->soft_joint_pos_limits || !a->prev_raw_actions ||
        !a->body_pos_w || !a->body_lin_vel_w || !a->net_forces_w_history || !a->current_air_time || !a->current_contact_time ||
        !a->last_air_time || !a->episode_length_buf || !a->reward || !a->terminated || !a->time_outs || !a->dones)
      return LT_ERR_INVALID_ARG;
    if (a->num_reward_terms <= 0 || a->num_reward_terms > LT_MAX_REWARD_TERMS || a->num_termination_terms < 0 ||
        a->num_termination_terms > LT_MAX_TERMINATION_TERMS || a->num_sensor_bodies <= 0 || a->num_sensor_bodies > kMaxSensorBodies ||
        a->force_history <= 0 || a->force_history > 8 || a->num_thigh_calf < 0 || a->num_thigh_calf > 8)
      return LT_ERR_INVALID_ARG;
    const LtGaitState& g = a->gait_state;
    if (!g.last_step_current_air_time || !g.last_step_current_contact_time || !g.swinging_in_zero_cmd || !g.valid_last_air_time ||
        !g.valid_previous_contact || !g.last_velocity_cmd || !g.step_from_changing_cmd)
      return LT_ERR_INVALID_ARG;
    bool seen[LT_RK_COUNT] = {false};
    for (int i = 0; i < a->num_reward_terms; ++i) {
      const int k = a->reward_terms[i].kind;
      if (k < 0 || k >= LT_RK_COUNT) return LT_ERR_INVALID_ARG;
      if (a->reward_terms[i].weight != 0.f) {
        if (seen[k]) return LT_ERR_UNSUPPORTED;  // one active term per kind
        seen[k] = true;
      }
      if (k >= LT_RK_OBJ_XY_POS && a->reward_terms[i].weight != 0.f && !has_obj) return LT_ERR_INVALID_ARG;
    }
    for (int t = 0; t < a->num_termination_terms; ++t) {
      const LtTerminationTerm& tt = a->termination_terms[t];
      if ((tt.kind == LT_TK_OBJECT_BELOW_ROBOT || tt.kind == LT_TK_BAD_ROLL) && !has_obj) return LT_ERR_INVALID_ARG;
      if (tt.num_ids < 0 || tt.num_ids > LT_MAX_CONTACT_IDS) return LT_ERR_INVALID_ARG;
    }
  }
  if (has_obj && (!a->root_quat_w || !a->root_lin_vel_w || !a->root_ang_vel_w || !a->obj_root_quat_w || !a->obj_root_lin_vel_w ||
                  !a->obj_root_ang_vel_w || !a->obj_projected_gravity_b))
    return LT_ERR_INVALID_ARG;
  Layout L;
  memset(&L, 0, sizeof(L));
  int dps = 0;
  if (do_obs) {
    if (a->num_obs_terms <= 0 || a->num_obs_terms > LT_MAX_OBS_TERMS || a->history_length <= 0 || !a->default_joint_vel)
      return LT_ERR_INVALID_ARG;
    if (!a->policy_obs_out && !a->critic_obs_out) return LT_ERR_INVALID_ARG;
    for (int t = 0; t < a->num_obs_terms; ++t) {
      const LtObsTerm& ot = a->obs_terms[t];
      if (ot.dim <= 0 || ot.dim > 255) return LT_ERR_INVALID_ARG;
      if (ot.kind == LT_OK_OBJECT_STATE && (!has_obj || ot.dim != 13)) return LT_ERR_INVALID_ARG;
      if ((ot.kind == LT_OK_JOINT_POS_REL || ot.kind == LT_OK_JOINT_VEL_REL || ot.kind == LT_OK_LAST_ACTION) && ot.dim != a->J)
        return LT_ERR_INVALID_ARG;
      if ((ot.kind == LT_OK_COMMAND || ot.kind == LT_OK_BASE_ANG_VEL || ot.kind == LT_OK_PROJECTED_GRAVITY) && ot.dim != 3)
        return LT_ERR_INVALID_ARG;
      dps += ot.dim;
    }
    L.D = dps * a->history_length;
    L.dps = dps;
    if (dps > kMaxNew || L.D > kMaxObsDim || dps > 255) return LT_ERR_UNSUPPORTED;
  }
  const int J = a->J, S = a->num_sensor_bodies > 0 ? a->num_sensor_bodies : 1, H = a->force_history > 0 ? a->force_history : 1;
  const int T = do_rew ? a->num_reward_terms : 0;
  int off = 0;
  auto take = [&](int per_env) { const int o = off; off += kEnvs * per_env; return o; };
  L.cmd = take(3); L.pos = take(3); L.linb = take(3); L.angb = take(3); L.grav = take(3);
  L.q = take(J); L.qd = take(J); L.qdd = take(J); L.tau = take(J); L.q0 = take(J); L.qd0 = take(J); L.lim = take(2 * J);
  L.act = take(J); L.pact = take(J);
  L.force = take(do_rew ? H * S * 3 : 0);
  L.air = take(4); L.con = take(4); L.lair = take(4); L.fpos = take(12); L.fvel = take(12);
  L.quat = take(4); L.linw = take(3); L.angw = take(3); L.opos = take(3); L.oquat = take(4); L.olin = take(3); L.oang = take(3);
  L.ograv = take(3); L.oc_last = take(1); L.oc_cur = take(1); L.oc_air = take(1);
  L.fmax = take(S);
  L.g_lsa = take(4); L.g_lsc = take(4); L.g_vla = take(4); L.g_cmd = take(3); L.g_steps = take(1); L.g_sz = take(4); L.g_vpc = take(4);
  L.eplen = take(2);  // int64 per env
  L.gait_out = take(kGaitFloats);
  L.esum = take(T); L.raw = take(T);
  L.newobs = take(2 * dps);
  L.uscr = take(16);
  off = (off + 3) & ~3;
  L.hist_stride = ((kEnvs * L.D + 3) & ~3) + 4;  // +4: the shifted read of the last element may touch one slot past the block
  L.hist = off; off += 2 * L.hist_stride;
  L.map = off; off += ((L.D + 3) & ~3) + dps;  // s_map[D] + s_jinfo[dps]
  L.total = off;
  const size_t smem = (size_t)L.total * sizeof(float);
  if (smem > 200 * 1024) return LT_ERR_UNSUPPORTED;

  StageTable ST;
  memset(&ST, 0, sizeof(ST));
  auto add = [&](const float* src, int row, int dst_off) {
    if (!src || row <= 0) return;
    ST.src[ST.n] = src; ST.row[ST.n] = row; ST.off[ST.n] = dst_off;
    ST.vec[ST.n] = ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (kEnvs * row) % 4 == 0 && (dst_off % 4) == 0) ? 1 : 0;
    ++ST.n;
  };
  if (do_rew) add(a->net_forces_w_history, H * S * 3, L.force);  // biggest first
  add(a->command, 3, L.cmd); add(a->root_pos_w, 3, L.pos); add(a->root_ang_vel_b, 3, L.angb); add(a->projected_gravity_b, 3, L.grav);
  add(a->joint_pos, J, L.q); add(a->joint_vel, J, L.qd); add(a->default_joint_pos, J, L.q0); add(a->raw_actions, J, L.act);
  if (do_obs) add(a->default_joint_vel, J, L.qd0);
  if (do_rew) {
    add(a->root_lin_vel_b, 3, L.linb); add(a->joint_acc, J, L.qdd); add(a->applied_torque, J, L.tau);
    add(a->soft_joint_pos_limits, 2 * J, L.lim); add(a->prev_raw_actions, J, L.pact);
  }
  if (do_rew) {
    const LtGaitState& gs = a->gait_state;
    add(gs.last_step_current_air_time, 4, L.g_lsa); add(gs.last_step_current_contact_time, 4, L.g_lsc); add(gs.valid_last_air_time, 4, L.g_vla);
    add(gs.last_velocity_cmd, 3, L.g_cmd); add(gs.step_from_changing_cmd, 1, L.g_steps);
  }
  if (has_obj) {
    if (!a->obj_last_contact_time || !a->obj_current_contact_time || !a->obj_current_air_time) return LT_ERR_INVALID_ARG;
    add(a->obj_last_contact_time, 1, L.oc_last); add(a->obj_current_contact_time, 1, L.oc_cur); add(a->obj_current_air_time, 1, L.oc_air);
    add(a->root_quat_w, 4, L.quat); add(a->root_lin_vel_w, 3, L.linw); add(a->root_ang_vel_w, 3, L.angw);
    add(a->obj_root_pos_w, 3, L.opos); add(a->obj_root_quat_w, 4, L.oquat); add(a->obj_root_lin_vel_w, 3, L.olin);
    add(a->obj_root_ang_vel_w, 3, L.oang); add(a->obj_projected_gravity_b, 3, L.ograv);
  }

  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mdp_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return lt::check(e);
    attr_set = true;
  }
  if (do_rew && a->any_nonzero_cmd_override == -2) {
    if (!a->any_flag_ws || a->offset_base) return LT_ERR_INVALID_ARG;
    any_nonzero_cmd_kernel<<<1, 1024, 0, st>>>(a->command, a->N, a->any_flag_ws, (int)a->offset);
    int rc = lt::check_launch();
    if (rc != LT_OK) return rc;
  }
  const int grid = (int)lt::ceil_div(a->N, kEnvs);
  LtMdpArgs args = *a;
  if (args.any_nonzero_cmd_override == -2) args.any_nonzero_cmd_override = -1;
  mdp_step_kernel<<<grid, kThreads, smem, st>>>(args, L, ST);
  return lt::check_launch();
}

extern "C" int lt_mdp_reset(const LtGaitState* g, float* episode_sums, int num_reward_terms, const uint8_t* mask, int N, void* stream) {
  if (!g || !mask || N <= 0) return LT_ERR_INVALID_ARG;
  mdp_reset_kernel<<<(unsigned)lt::ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(*g, episode_sums, num_reward_terms, mask, N);
  return lt::check_launch();
}
