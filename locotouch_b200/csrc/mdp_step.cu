// K1: fused MDP step -- terminations -> rewards (incl. the stateful adaptive symmetric gait reward) -> [auto reset]
// -> policy / critic observations with history shift and noise.
// Replaces the per-term PyTorch code of reference locotouch/mdp/rewards.py:15-604, terminations.py:10-23,
// observations.py:38-91 and the IsaacLab manager loops around it (several thousand ATen launches per env step,
// SURVEY.md 2.1) with one launch that reads exactly the IsaacLab tensors the reference terms read.
//
// Block = 1024 threads = 32 consecutive envs: LANE == ENV in every per-env computation, WARP == TASK.  The rows of 32
// consecutive envs are ONE contiguous span of every [N, row] input tensor, so all global traffic is coalesced and the
// per-env arithmetic runs out of shared memory.  Every task is a serial dependent chain (measured: ~5 us each, cold or warm),
// so the kernel is organised to run as many short chains side by side as an SM holds (32 warps, one block per SM).
//   stage 0  one bulk async copy (cp.async.bulk, the TMA 1-D path; mbarrier completion) per state tensor, issued by the first
//            lanes of every warp; gathers (feet rows, [terms, N] episode sums) one 4-byte cp.async per thread; launch-constant
//            lookup tables arrive prebuilt (lt_mdp_build_tables).  The two observation history blocks (the bulk of the bytes)
//            are requested after the state has landed and arrive while stage 1 runs.  Tail block / unaligned callers: per
//            element cp.async.
//   stage 1  18 fixed tasks + one task per 4 new observation values; first round task == warp, then a shared queue:
//              0,1    gait state update + swing bonus of one synced pair            rewards.py:158-346
//              2      gait task score, async / stance parts, final gait term        rewards.py:202-216, 348-392
//              3,4    terminations (even / odd terms), alive                        terminations.py:10-23, rewards.py:15-22
//              5-10   robot reward terms                                            rewards.py:24-56, 398-466
//              11-13  object-transport reward terms                                 rewards.py:469-604
//              14-17  object_state_in_robot_frame: policy pos+vel / ang-vel / quat (noise), critic   observations.py:38-91
//              18+    proprioceptive observation values (noise, scale)              [IL] ObservationManager
//   stage 2  all threads: weighted accumulation in manager order, episode sums, per-term outputs, gait state write back
//            (zeroed for reset envs), and the history shift: shared -> global, one warp per env row, 8-byte stores.
// Floating point follows the reference's fp32 expression order (file is built with -fmad=false); masks are bit-exact.
#include <cuda_pipeline.h>
#include <string.h>

#include "lt_common.cuh"

namespace {

constexpr int kEnvs = 32;  // envs per block == lanes per warp
constexpr int kThreads = 1024;  // 32 warps: the per-warp work is a serial dependent chain, so more (shorter) chains per SM
                                          // is what hides latency; one block per SM
constexpr int kWarps = kThreads / 32;
constexpr int kMaxObsDim = 512;
constexpr int kMaxNew = 96;
constexpr int kMaxJ = 16;
constexpr int kMaxSensorBodies = 32;
constexpr int kMaxStage = 36;
constexpr int kGaitFloats = 24;  // per env: lsa[4] lsc[4] vla[4] last_cmd[3] steps sz[4] vpc[4]
constexpr int kGaitStride = 25;
constexpr int kFixedTasks = 18;
constexpr int kS3 = 3, kS4 = 4;    // per-env strides of the bulk-copied 3- and 4-float rows (= row length)
constexpr int kP4 = 5, kS12 = 13;  // padded strides of the rows the kernel gathers element by element

// shared-memory layout (float offsets), filled in on the host.  Bulk-copied rows keep their row length as per-env stride.
struct Layout {
  int cmd, pos, linb, angb, grav, q, qd, qdd, tau, q0, qd0, lim, act, pact, force, air, con, lair, fpos, fvel;
  int quat, linw, angw, opos, oquat, olin, oang, ograv, oc_last, oc_cur, oc_air, g_lsa, g_lsc, g_vla, g_cmd, g_steps, g_sz, g_vpc, eplen, gait_out,
      esum, raw, gtmp, fmax, oframe, newobs, hist, map, total;
  int sJ, sLim, sF;  // strides of the [J], [J, 2] and [H, S, 3] rows
  int D;             // observation dim per group
  int dps;           // new values per step per group
  int sNew;          // stride of the new-value rows (dps | 1)
  int hist_stride;   // floats per group in the history staging area (kEnvs * D rounded up to 4, + 4)
  int tables_len;    // ints in the table block (see tables_len())
  int bulk_ok;       // every [N, row] tensor base is 16-byte aligned: full blocks stage with cp.async.bulk
  int hist_bulk_ok;  // likewise for the two history sources
  unsigned tx_state; // bytes the state barrier of a full block waits for
};

// global -> shared copy descriptors for the plain [N, row] float tensors (the block's kEnvs rows are one contiguous span)
struct StageTable {
  int n;
  const float* src[kMaxStage];
  int off[kMaxStage];
  int row[kMaxStage];
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
// one shared copy of the generator: the kernel is bound by instruction fetch (every warp runs different code), so code size matters
__device__ __noinline__ uint4 philox4(uint64_t seed, uint64_t offset, uint32_t a, uint32_t b) { return lt::Philox::gen(seed, offset, a, b); }

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

// Optional in-kernel timeline (profiling builds only: LT_MDP_PROF=1 python -m locotouch_b200.csrc.build --force; tools/mdp_timeline.py)
#ifdef LT_MDP_PROF
__device__ long long* g_mdp_prof = nullptr;  // [blocks][kWarps][32] clock64 stamps
#define PROF_STAMP(slot)                                                                                   \
  do {                                                                                                     \
    if (g_mdp_prof && lane == 0) g_mdp_prof[((size_t)blockIdx.x * kWarps + warp) * 32 + (slot)] = clock64(); \
  } while (0)
#else
#define PROF_STAMP(slot) do { } while (0)
#endif

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// global -> shared bulk async copy (TMA 1-D): 16-byte aligned addresses, size a multiple of 16; completion on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void named_barrier(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// max over the history of |F| of one sensor body: torch.max(torch.norm(net_forces_w_history[:, :, ids], dim=-1), dim=1)[0]
// The H norms are independent chains (three in flight for the usual H == 3); the max is order-independent.
__device__ __forceinline__ float body_force_max(const float* force_env, int H, int S, int b) {
  float m = 0.f;
  int h = 0;
  for (; h + 3 <= H; h += 3) {
    const float* p0 = force_env + (h * S + b) * 3;
    const float* p1 = p0 + S * 3;
    const float* p2 = p1 + S * 3;
    const float n0 = sqrtf(p0[0] * p0[0] + p0[1] * p0[1] + p0[2] * p0[2]);
    const float n1 = sqrtf(p1[0] * p1[0] + p1[1] * p1[1] + p1[2] * p1[2]);
    const float n2 = sqrtf(p2[0] * p2[0] + p2[1] * p2[1] + p2[2] * p2[2]);
    m = fmaxf(m, fmaxf(n0, fmaxf(n1, n2)));
  }
#pragma unroll 1
  for (; h < H; ++h) {
    const float* p = force_env + (h * S + b) * 3;
    m = fmaxf(m, sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
  }
  return m;
}

// Measured and rejected (round 1, tools/mdp_timeline.py): touching every 64-byte line of the 3.4 KB parameter block up front, or
// copying the block into shared memory, changes nothing (15.08 vs 15.06 us) or costs 1 us -- parameter reads are not what the
// chains wait for; writing the old 5/6 of every observation row from the task queue (ahead of stage 2b) only moves the same
// latency-bound copy loop earlier (14.3 vs 13.9 us).  What did pay: see the reward sum in stage 2a.
__global__ void __launch_bounds__(kThreads, 1) mdp_step_kernel(const LtMdpArgs A, const Layout L, const StageTable ST) {
  extern __shared__ __align__(16) float sm[];
  __shared__ unsigned char s_done[kEnvs], s_fill[kEnvs], s_tflag[kEnvs], s_tout[kEnvs];
  __shared__ int s_step, s_any_nz, s_next_task;
  __shared__ __align__(8) unsigned long long s_bar[2];  // mbarriers: [0] state tensors + tables, [1] observation history
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e0 = blockIdx.x * kEnvs;
  const int nvalid = min(kEnvs, A.N - e0);
  const bool do_rew = A.phases & LT_PHASE_REWARDS, do_obs = A.phases & LT_PHASE_OBS;
  const bool has_obj = A.obj_root_pos_w != nullptr;
  const int J = A.J, S = A.num_sensor_bodies, H = A.force_history;
  const int D = L.D, dps = L.dps, T = A.num_reward_terms;
  const int sJ = L.sJ, sNew = L.sNew;

  PROF_STAMP(0);
  // ------------------------------------------------------------------------------------------------ stage 0: loads
  // Full blocks with 16-byte aligned tensors: ONE bulk async copy (cp.async.bulk, the TMA 1-D path) per tensor, issued by
  // the lanes of warp 0 and tracked by an mbarrier -- a few dozen instructions instead of one cp.async per element.  The tail
  // block (nvalid < kEnvs) and unaligned callers take the per-element path below.
  const bool bulk = L.bulk_ok && nvalid == kEnvs;
  const bool bulk_hist = do_obs && L.hist_bulk_ok && nvalid == kEnvs;
  int* s_map = reinterpret_cast<int*>(sm + L.map);
  int* s_jinfo = s_map + ((D + 3) & ~3);
  int* s_slot = s_jinfo + dps;  // reward kind -> index in the term table (-1: absent or zero weight)
  const float* s_kpar = reinterpret_cast<const float*>(s_slot + LT_RK_COUNT);  // [kind][6] parameters of the active term of each kind
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    s_next_task = kWarps;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // every warp issues the copies of tensors warp, warp + 16, ... from its first lanes (a bulk copy is one instruction per lane);
  // thread 0 posts the expected byte count -- the barrier cannot complete before that arrival, whatever order the copies finish in
  if (bulk) {
    if (tid == 0) mbar_expect_tx(&s_bar[0], L.tx_state);
    const int t = warp + kWarps * lane;
    if (t < ST.n) bulk_g2s(sm + ST.off[t], ST.src[t] + (size_t)e0 * ST.row[t], (unsigned)(kEnvs * 4) * ST.row[t], &s_bar[0]);
    if (tid == kThreads - 32 && A.tables) bulk_g2s(s_map, A.tables, (unsigned)L.tables_len * 4u, &s_bar[0]);
  }
  if (!bulk) {
#pragma unroll 1
    for (int t = 0; t < ST.n; ++t) {
      const float* src = ST.src[t] + (size_t)e0 * ST.row[t];
      float* dst = sm + ST.off[t];
      const int count = nvalid * ST.row[t];
#pragma unroll 1
      for (int i = tid; i < count; i += kThreads) __pipeline_memcpy_async(dst + i, src + i, 4);
    }
    if (A.tables) {
#pragma unroll 1
      for (int i = tid; i < L.tables_len; i += kThreads) s_map[i] = A.tables[i];
    }
  }
  PROF_STAMP(16);
  // gathers with arbitrary body ids / non-float types: ONE pass, every thread issues a few independent copies
  if (do_rew) {
    if (tid >= 384 && tid < 384 + nvalid * 4) {  // feet rows of the contact timers
      const int i = tid - 384, e = i >> 2, k = i & 3;
      const size_t r = (size_t)(e0 + e) * S + A.gait.feet_ids[k];
      __pipeline_memcpy_async(sm + L.air + e * kP4 + k, A.current_air_time + r, 4);
      __pipeline_memcpy_async(sm + L.con + e * kP4 + k, A.current_contact_time + r, 4);
      __pipeline_memcpy_async(sm + L.lair + e * kP4 + k, A.last_air_time + r, 4);
    }
    if (tid < nvalid * 12) {  // feet rows of body_pos_w / body_lin_vel_w
      const int e = tid / 12, rc = tid - e * 12, k = rc / 3, c = rc - k * 3;
      const size_t r = ((size_t)(e0 + e) * A.num_bodies + A.feet_body_ids[k]) * 3 + c;
      __pipeline_memcpy_async(sm + L.fpos + e * kS12 + rc, A.body_pos_w + r, 4);
      __pipeline_memcpy_async(sm + L.fvel + e * kS12 + rc, A.body_lin_vel_w + r, 4);
    }
    if (A.episode_sums) {  // [terms][N]: kEnvs contiguous floats per term
#pragma unroll 1
      for (int i = tid; i < T * kEnvs; i += kThreads) {
        const int t = i >> 5, e = i & 31;
        if (e < nvalid) __pipeline_memcpy_async(sm + L.esum + i, A.episode_sums + (size_t)t * A.N + e0 + e, 4);
      }
    }
  }
  if (tid == kThreads - 1) {  // env-step index and the cross-env any(non_zero_cmd) flag (one dependent pair of loads, one thread)
    const int st = (int)(A.offset + (A.offset_base ? (uint64_t)*A.offset_base : 0ull));
    s_step = st;
    int any_nz = 1;
    if (A.any_nonzero_cmd_override >= 0) any_nz = A.any_nonzero_cmd_override != 0;
    else if (A.any_flag_ws) any_nz = A.any_flag_ws[st & 1] == st;
    s_any_nz = any_nz;
  }
  PROF_STAMP(17);
  __pipeline_commit();
  if (do_obs && !bulk_hist) {  // observation history blocks, per-element path
    const int total = nvalid * D;
#pragma unroll 1
    for (int grp = 0; grp < 2; ++grp) {
      const float* in = grp ? A.critic_obs_in : A.policy_obs_in;
      if (!in) continue;
      const float* src = in + (size_t)e0 * D;
      float* dst = sm + L.hist + grp * L.hist_stride;
#pragma unroll 1
      for (int i = tid; i < total; i += kThreads) __pipeline_memcpy_async(dst + i, src + i, 4);
    }
  }
  __pipeline_commit();
  // non-float state: plain loads, issued after every asynchronous copy is in flight
  if (do_rew) {
    if (tid < nvalid * 4) {
      const float sz = (float)A.gait_state.swinging_in_zero_cmd[(size_t)e0 * 4 + tid];
      const float vpc = (float)A.gait_state.valid_previous_contact[(size_t)e0 * 4 + tid];
      sm[L.g_sz + (tid >> 2) * kP4 + (tid & 3)] = sz;
      sm[L.g_vpc + (tid >> 2) * kP4 + (tid & 3)] = vpc;
    }
    if (tid >= 128 && tid < 128 + nvalid) reinterpret_cast<long long*>(sm + L.eplen)[tid - 128] = A.episode_length_buf[e0 + tid - 128];
  }
  if (tid >= 160 && tid < 160 + kEnvs) {
    const int i = tid - 160;
    s_fill[i] = (i < nvalid && do_obs && A.obs_fill) ? A.obs_fill[e0 + i] : 0;
    s_done[i] = 0;
    s_tout[i] = 0;
  }
  PROF_STAMP(18);
  if (!A.tables) {
    // Observation layout tables built in the kernel (callers that do not pass lt_mdp_build_tables() output):
    //   s_map[k]  for column k of the flattened [term][history][dim] row: low 16 bits = index j of the per-step value that
    //             feeds the column (used when the history is (re)filled), high 16 bits = k + d, the column one history
    //             slot later of the same term (the shift source), or 0xffff when k is the newest slot
    //   s_jinfo[j] for per-step value j: term index | component << 8
    if (do_obs) {
#pragma unroll 1
      for (int k = tid; k < D + dps; k += kThreads) {
        int col = 0, jbase = 0;
        if (k < D) {
          int entry = 0;
#pragma unroll 1
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
    if (tid >= 256 && tid < 256 + LT_RK_COUNT) {
      const int kind = tid - 256;
      int slot = -1;
#pragma unroll 1
      for (int i = 0; i < T; ++i)
        if (A.reward_terms[i].kind == kind && A.reward_terms[i].weight != 0.f) slot = i;
      s_slot[kind] = slot;
#pragma unroll 1
      for (int k = 0; k < 6; ++k) reinterpret_cast<float*>(s_slot + LT_RK_COUNT)[kind * 6 + k] = slot >= 0 ? A.reward_terms[slot].p[k] : 0.f;
    }
  }
  PROF_STAMP(1);
  __pipeline_wait_prior(1);  // gathers (and the per-element state copies) have landed; the history may still be in flight
  if (bulk) mbar_wait(&s_bar[0], 0);
  PROF_STAMP(2);
  __syncthreads();
  PROF_STAMP(3);
  // the history blocks (the bulk of the bytes) are requested only now: they are not needed before stage 2 and would otherwise
  // compete with the state tensors for the bandwidth of the very first microseconds
  if (bulk_hist && warp == kWarps - 2) {
    const unsigned bytes = (unsigned)(kEnvs * 4) * D;
    if (lane == 0) mbar_expect_tx(&s_bar[1], (A.policy_obs_in ? bytes : 0u) + (A.critic_obs_in ? bytes : 0u));
    if (lane < 2) {
      const float* in = lane ? A.critic_obs_in : A.policy_obs_in;
      if (in) bulk_g2s(sm + L.hist + lane * L.hist_stride, in + (size_t)e0 * D, bytes, &s_bar[1]);
    }
  }
  const int step = s_step;
  const uint64_t rng_offset = (uint64_t)(int64_t)step;
  // Wide pre-phase (warps run the same short code side by side; everything below only combines these values):
  //  * contact-force maxima of every sensor body, one warp per body -- the terminations and the slip / collision terms compare
  //    them against their thresholds;
  //  * the object's state in the robot frame (rewards.py:469-543, observations.py:55-58), five warps: position, linear velocity,
  //    angular velocity, gravity direction, relative quaternion -- shared by the object reward terms and both observation groups.
  float* s_of = sm + L.oframe;  // [16][kEnvs]
  if (do_rew && lane < nvalid) {
#pragma unroll 1
    for (int b = warp; b < S; b += kWarps) sm[L.fmax + b * kEnvs + lane] = body_force_max(sm + L.force + lane * L.sF, H, S, b);
  }
  if (has_obj && warp >= kWarps - 5 && lane < nvalid) {
    const int job = kWarps - 1 - warp, e = lane;
    const Quat q = ld4(sm + L.quat + e * kS4);
    if (job < 4) {
      Vec3 d;
      if (job == 0) d = sub3(ld3(sm + L.opos + e * kS3), ld3(sm + L.pos + e * kS3));
      else if (job == 1) d = sub3(ld3(sm + L.olin + e * kS3), ld3(sm + L.linw + e * kS3));
      else if (job == 2) d = sub3(ld3(sm + L.oang + e * kS3), ld3(sm + L.angw + e * kS3));
      else d = rot(ld4(sm + L.oquat + e * kS4), ld3(sm + L.ograv + e * kS3));
      const Vec3 r = rot_inv(q, d);
      s_of[(3 * job + 0) * kEnvs + e] = r.x; s_of[(3 * job + 1) * kEnvs + e] = r.y; s_of[(3 * job + 2) * kEnvs + e] = r.z;
    } else {
      const Quat qr = quat_mul(quat_inv(q), ld4(sm + L.oquat + e * kS4));
      s_of[12 * kEnvs + e] = qr.w; s_of[13 * kEnvs + e] = qr.x; s_of[14 * kEnvs + e] = qr.y; s_of[15 * kEnvs + e] = qr.z;
    }
  }
  // Fused action term (K0 inside K1): JointPositionActionPrevPrev.process_actions on the staged action-term state of the block's envs,
  // the arithmetic of lt_process_actions bit for bit (explicitly rounded multiply / add); the tasks below read the updated rows.
  if (A.act_new) {
#pragma unroll 1
    for (int i = tid; i < nvalid * J; i += kThreads) {
      const int ee = i / J, c = i - ee * J;
      const size_t g = (size_t)(e0 + ee) * J + c;
      const float old_raw = sm[L.act + ee * sJ + c], old_prev = sm[L.pact + ee * sJ + c];
      float a = A.act_new[g];
      if (A.act_clip > 0.f) a = fminf(fmaxf(a, -A.act_clip), A.act_clip);
      a = __fmul_rn(a, A.act_raw_scale);
      if (A.act_prev_prev_raw) A.act_prev_prev_raw[g] = old_prev;
      const_cast<float*>(A.prev_raw_actions)[g] = old_raw;
      const_cast<float*>(A.raw_actions)[g] = a;
      if (A.act_processed) A.act_processed[g] = __fadd_rn(__fmul_rn(a, A.act_scale), A.act_offset ? A.act_offset[g] : 0.f);
      sm[L.act + ee * sJ + c] = a;
      sm[L.pact + ee * sJ + c] = old_raw;
    }
  }
  PROF_STAMP(21);
  if (do_rew || has_obj || A.act_new) __syncthreads();

  // ------------------------------------------------------------------------------------------------ stage 1: tasks
  float* s_raw = sm + L.raw;  // [LT_RK_COUNT][kEnvs]: unweighted value of every reward kind (stage 2a picks the task's terms)
  float* nv = sm + L.newobs;  // [2 groups][kEnvs][sNew]
  const int e = lane, n = e0 + e;
  const bool live = e < nvalid;
  const int quads = do_obs ? (dps + 3) >> 2 : 0;
  auto put = [&](int kind, float v) { s_raw[kind * kEnvs + e] = v; };
  auto par = [&](int kind, int k) -> float { return s_kpar[kind * 6 + k]; };
  // publish any(non_zero_cmd) for the reward pass of the NEXT step (it sees the same command tensor)
  if (do_obs && A.any_flag_ws && warp == kWarps - 1 && live) {
    const float* c = sm + L.cmd + e * kS3;
    if (sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) > 0.f) A.any_flag_ws[(step + 1) & 1] = step + 1;
  }
  // first round static (task == warp: the three gait tasks meet at a named barrier), then a shared queue: warps whose fixed
  // task was short pick up the per-quad observation tasks while the long ones are still running
  auto next_task = [&]() -> int {
    int t = 0;
    if (lane == 0) t = atomicAdd(&s_next_task, 1);
    return __shfl_sync(LT_FULL_MASK, t, 0);
  };
#ifdef LT_MDP_TWICE  // experiment: run the (idempotent) task phase twice; the second pass sees warm instruction / constant caches
  for (int pass = 0; pass < 2; ++pass) {
  if (pass) {
    __syncthreads();
    if (tid == 0) s_next_task = kWarps;
    __syncthreads();
    PROF_STAMP(12);
  }
#endif
#pragma unroll 1
  for (int task = warp; task < kFixedTasks + quads; task = next_task()) {
    PROF_STAMP(task < kWarps ? 8 : 10);
#ifdef LT_MDP_ONETASK  // experiment: every warp runs the same (barrier-free) task -> isolates instruction-supply effects
    if (task >= kWarps) continue;
    task = LT_MDP_ONETASK;
#endif
    if (task <= 2) {
      // ---------------------------------------------------------------------------------------------- gait reward
      if (!do_rew) continue;
      const int gi = s_slot[LT_RK_GAIT];
      if (gi < 0) {  // no gait term: carry the state through unchanged
        if (task == 0 && live) {
          float* go = sm + L.gait_out + e * kGaitStride;
#pragma unroll 1
          for (int k = 0; k < 4; ++k) {
            go[k] = sm[L.g_lsa + e * kS4 + k]; go[4 + k] = sm[L.g_lsc + e * kS4 + k]; go[8 + k] = sm[L.g_vla + e * kS4 + k];
            go[16 + k] = sm[L.g_sz + e * kP4 + k]; go[20 + k] = sm[L.g_vpc + e * kP4 + k];
          }
          go[12] = sm[L.g_cmd + e * kS3]; go[13] = sm[L.g_cmd + e * kS3 + 1]; go[14] = sm[L.g_cmd + e * kS3 + 2]; go[15] = sm[L.g_steps + e];
        }
        continue;
      }
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
      float* gtmp = sm + L.gtmp;  // [pair][bonus, code][kEnvs]
      const Vec3 cmd = ld3(sm + L.cmd + e * kS3);
      const bool nz = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z) > 0.f;
      float a[4], c[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { a[k] = sm[L.air + e * kP4 + k]; c[k] = sm[L.con + e * kP4 + k]; }
      if (task < 2) {
        // tasks 0 / 1: state update (rewards.py:158-200; both warps compute it, task 0 stores it) + one synced pair (:218-241)
        if (live) {
          GaitRegs g;
          float la[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            g.lsa[k] = sm[L.g_lsa + e * kS4 + k]; g.lsc[k] = sm[L.g_lsc + e * kS4 + k]; g.vla[k] = sm[L.g_vla + e * kS4 + k];
            g.sz[k] = sm[L.g_sz + e * kP4 + k] != 0.f; g.vpc[k] = sm[L.g_vpc + e * kP4 + k] != 0.f;
            la[k] = sm[L.lair + e * kP4 + k];
          }
          g.last_cmd[0] = sm[L.g_cmd + e * kS3]; g.last_cmd[1] = sm[L.g_cmd + e * kS3 + 1]; g.last_cmd[2] = sm[L.g_cmd + e * kS3 + 2];
          g.steps = sm[L.g_steps + e];
          gait_update(g, a, c, la, cmd, nz, s_any_nz != 0, th);
          if (task == 0) {
            float* go = sm + L.gait_out + e * kGaitStride;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              go[k] = g.lsa[k]; go[4 + k] = g.lsc[k]; go[8 + k] = g.vla[k];
              go[16 + k] = g.sz[k] ? 1.f : 0.f; go[20 + k] = g.vpc[k] ? 1.f : 0.f;
            }
            go[12] = g.last_cmd[0]; go[13] = g.last_cmd[1]; go[14] = g.last_cmd[2]; go[15] = g.steps;
          }
          // pair 0 = gait feet (0, 1), pair 1 = (2, 3); selects instead of indexing keep the arrays in registers
          const float af0 = task ? a[2] : a[0], af1 = task ? a[3] : a[1], cf0 = task ? c[2] : c[0], cf1 = task ? c[3] : c[1];
          const float vt0 = task ? g.vla[2] : g.vla[0], vt1 = task ? g.vla[3] : g.vla[1];
          const float vo0 = task ? g.vla[0] : g.vla[2], vo1 = task ? g.vla[1] : g.vla[3];
          const bool both_air = af0 > th && af0 < gp.air_time_gait_bound && af1 > th && af1 < gp.air_time_gait_bound;
          const bool c0 = cf0 > th && cf0 < gp.contact_time_gait_bound;
          const bool c1 = cf1 > th && cf1 < gp.contact_time_gait_bound;
          float bonus = 0.f;
          if (gp.encourage_symmetricity) bonus = gait_swing_bonus(gp, af0, af1, vt0, vt1, vo0, vo1);
          gtmp[(task * 2 + 0) * kEnvs + e] = bonus;
          gtmp[(task * 2 + 1) * kEnvs + e] = both_air ? 2.f : ((c0 && c1) ? 1.f : 0.f);
        }
        named_barrier(1, 96);
      } else {
        // task 2: task score (rewards.py:202-216 / 372-392), async pairs (:348-363), stance; then the final combination
        float score = 0.f, asyn = 0.f, stance = 0.f;
        if (live) {
          if (gp.encourage_symmetricity) {
            const Vec3 vb = ld3(sm + L.linb + e * kS3);
            const float wz = sm[L.angb + e * kS3 + 2];
            const float dx = cmd.x - vb.x, dy = cmd.y - vb.y;
            const float e_lin = nz ? sqrtf(dx * dx + dy * dy) : 0.f;
            const float e_ang = nz ? fabsf(cmd.z - wz) : 0.f;
            score = (exp_neg_over(e_lin, A.gait.vel_tracking_exp_sigma) + exp_neg_over(e_ang, A.gait.vel_tracking_exp_sigma)) / 2.f;
            if (A.gait.with_object) {
              const Vec3 rel_w = sub3(ld3(sm + L.opos + e * kS3), ld3(sm + L.pos + e * kS3));
              const Vec3 r = rot_inv(yaw_quat(yaw_of(ld4(sm + L.quat + e * kS4))), rel_w);
              const float bx = clampf(1.f - fabsf(r.x) / A.gait.obj_x_max, 0.f, 1.f);
              const float by = clampf(1.f - fabsf(r.y) / A.gait.obj_y_max, 0.f, 1.f);
              score = clampf((score * 2.f + (bx + by) / 2.f) / 3.f, 0.f, 1.f);
            }
          }
          asyn = (gait_async(gp, a[0], a[2], c[0], c[2]) + gait_async(gp, a[1], a[3], c[1], c[3]) +
                  gait_async(gp, a[0], a[3], c[0], c[3]) + gait_async(gp, a[2], a[1], c[2], c[1])) / 4.f;
          stance = ((c[0] > th && c[1] > th && c[2] > th && c[3] > th) ? 1.f : 0.f) * A.gait.stance_rwd_scale;
        }
        named_barrier(1, 96);
        if (live) {
          float v[2];
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const float code = gtmp[(p * 2 + 1) * kEnvs + e];
            if (gp.encourage_symmetricity) {
              float bonus = gtmp[(p * 2 + 0) * kEnvs + e];
              const float scale = 1.f - gp.task_performance_ratio + gp.task_performance_ratio * score;
              if (bonus > 0.f) bonus *= scale;
              bonus += 1.f;
              v[p] = code == 2.f ? bonus : (code == 1.f ? 1.f : 0.f);
            } else {
              v[p] = code != 0.f ? 1.f : 0.f;
            }
          }
          const float sync = (v[0] + v[1]) / 2.f;
          const float stepping = (sync + asyn) / 2.f;
          s_raw[LT_RK_GAIT * kEnvs + e] = nz ? stepping : stance;
        }
      }
      continue;
    }
    if (task >= kFixedTasks) {
      // -------------------------------------------------- new observation values of the proprioceptive terms, 4 per task
      if (!live) continue;
      const int qd = task - kFixedTasks;
      uint4 rnd = make_uint4(0, 0, 0, 0);
      if (!A.u_obs) rnd = philox4(A.seed, rng_offset, (uint32_t)n, (uint32_t)qd);
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
          case LT_OK_COMMAND: raw = sm[L.cmd + e * kS3 + c]; break;
          case LT_OK_BASE_ANG_VEL: raw = sm[L.angb + e * kS3 + c]; break;
          case LT_OK_PROJECTED_GRAVITY: raw = sm[L.grav + e * kS3 + c]; break;
          case LT_OK_JOINT_POS_REL: raw = sm[L.q + e * sJ + c] - sm[L.q0 + e * sJ + c]; break;
          case LT_OK_JOINT_VEL_REL: raw = sm[L.qd + e * sJ + c] - sm[L.qd0 + e * sJ + c]; break;
          default: raw = sm[L.act + e * sJ + c]; break;  // LT_OK_LAST_ACTION
        }
        float noisy = raw;
        if (ot.noisy) {
          const float u = A.u_obs ? __ldcs(A.u_obs + (size_t)n * dps + j) : lt::Philox::u01(rw[c4]);
          noisy = (raw + u * (ot.n_max - ot.n_min)) + ot.n_min;  // [IL] data + rand*(max-min) + min
        }
        nv[(0 * kEnvs + e) * sNew + j] = noisy * ot.scale;
        nv[(1 * kEnvs + e) * sNew + j] = raw * ot.scale;
      }
      continue;
    }
    if (task >= 14) {
      // ------------------------------------------- object_state_in_robot_frame (observations.py:38-91)
      // 14: policy group, position + linear velocity slots; 15: policy group, angular velocity; 16: policy group, quaternion;
      // 17: critic group (all slots, no noise)
      if (!do_obs || !has_obj) continue;
      int jb = 0, tobj = -1;
#pragma unroll 1
      for (int t = 0; t < A.num_obs_terms; ++t) {
        if (A.obs_terms[t].kind == LT_OK_OBJECT_STATE) { tobj = t; break; }
        jb += A.obs_terms[t].dim;
      }
      if (tobj < 0 || !live) continue;
      const int grp = task == 17 ? 1 : 0;
      const bool do_pv = task == 14 || task == 17, do_w = task == 15 || task == 17, do_q = task == 16 || task == 17;
      const Quat q = ld4(sm + L.quat + e * kS4);
      const bool never = sm[L.oc_last + e] < A.os_last_contact_thr && sm[L.oc_cur + e] < A.os_current_contact_thr;
      const float term_scale = A.obs_terms[tobj].scale;
      float* out = nv + (grp * kEnvs + e) * sNew + jb;
      const bool noise = grp == 0;  // add_uniform_noise=True for the policy group only
      if (do_pv || do_w) {
        // position / linear velocity / angular velocity slots: raw values first, then one pass per Philox quad
        if (do_pv) {
          const Vec3 p = {s_of[0 * kEnvs + e], s_of[1 * kEnvs + e], s_of[2 * kEnvs + e]};
          const Vec3 v = {s_of[3 * kEnvs + e], s_of[4 * kEnvs + e], s_of[5 * kEnvs + e]};
          out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = v.x; out[4] = v.y; out[5] = v.z;
        }
        if (do_w) {
          const Vec3 w = {s_of[6 * kEnvs + e], s_of[7 * kEnvs + e], s_of[8 * kEnvs + e]};
          out[10] = w.x; out[11] = w.y; out[12] = w.z;
        }
#pragma unroll 1
        for (int qd = jb >> 2; qd <= (jb + 12) >> 2; ++qd) {
          const int k0 = 4 * qd - jb;  // slots k0 .. k0+3 of this quad
          const bool any_mine = (do_pv && k0 <= 5 && k0 + 3 >= 0) || (do_w && k0 <= 12 && k0 + 3 >= 10);
          if (!any_mine) continue;
          uint4 r4 = make_uint4(0, 0, 0, 0);
          if (noise && !A.u_obs) r4 = philox4(A.seed, rng_offset, (uint32_t)n, (uint32_t)qd);
          const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int k = k0 + c4;
            const bool mine = (do_pv && k >= 0 && k <= 5) || (do_w && k >= 10 && k <= 12);
            if (!mine) continue;
            float st = out[k], cst = A.os_non_contact[k];
            if (noise) {
              const float u = A.u_obs ? __ldcs(A.u_obs + (size_t)n * dps + jb + k) : lt::Philox::u01(rw[c4]);
              const float add = u * (A.os_n_max[k] - A.os_n_min[k]) + A.os_n_min[k];  // observations.py:77
              st = st + add;
              cst = cst + add;  // observations.py:82 (same draw)
            }
            out[k] = ((never ? cst : st) * A.os_scale[k]) * term_scale;
          }
        }
      }
      if (do_q) {
        const Quat qr = {s_of[12 * kEnvs + e], s_of[13 * kEnvs + e], s_of[14 * kEnvs + e], s_of[15 * kEnvs + e]};
        float st[4] = {qr.w, qr.x, qr.y, qr.z};
        float cst[4] = {A.os_non_contact[6], A.os_non_contact[7], A.os_non_contact[8], A.os_non_contact[9]};
        if (noise) {  // additive slots, then euler-angle noise composed onto the quaternion
          float u[4];
          if (A.u_obs) {
#pragma unroll
            for (int k = 0; k < 4; ++k) u[k] = __ldcs(A.u_obs + (size_t)n * dps + jb + 6 + k);
          } else {  // value j takes word j & 3 of quad j >> 2: the four quaternion slots straddle at most two quads
            const int j0 = jb + 6;
            const uint4 ra = philox4(A.seed, rng_offset, (uint32_t)n, (uint32_t)(j0 >> 2));
            uint4 rb = ra;
            if (j0 & 3) rb = philox4(A.seed, rng_offset, (uint32_t)n, (uint32_t)((j0 >> 2) + 1));
            const uint32_t w8[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
            const int sh = j0 & 3;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t wv = sh == 0 ? w8[k] : (sh == 1 ? w8[k + 1] : (sh == 2 ? w8[k + 2] : w8[k + 3]));
              u[k] = lt::Philox::u01(wv);
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float add = u[k] * (A.os_n_max[6 + k] - A.os_n_min[6 + k]) + A.os_n_min[6 + k];
            st[k] = st[k] + add;
            cst[k] = cst[k] + add;
          }
          float de[3];
          uint4 er = make_uint4(0, 0, 0, 0);
          if (!A.u_obj_euler) er = philox4(A.seed, rng_offset, (uint32_t)n, 0x1000u >> 2);  // values 0x1000 + {0,1,2}
          const uint32_t ew[3] = {er.x, er.y, er.z};
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float uu = A.u_obj_euler ? A.u_obj_euler[n * 3 + k] : lt::Philox::u01(ew[k]);
            de[k] = uu * (A.os_euler_max[k] - A.os_euler_min[k]) + A.os_euler_min[k];
          }
          const Quat nq = quat_from_euler(de[0], de[1], de[2]);  // observations.py:78-79
          const Quat a = quat_mul({st[0], st[1], st[2], st[3]}, nq);
          st[0] = a.w; st[1] = a.x; st[2] = a.y; st[3] = a.z;
          const Quat b = quat_mul({cst[0], cst[1], cst[2], cst[3]}, nq);
          cst[0] = b.w; cst[1] = b.x; cst[2] = b.y; cst[3] = b.z;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) out[6 + k] = ((never ? cst[k] : st[k]) * A.os_scale[6 + k]) * term_scale;
      }
      continue;
    }
    // ------------------------------------------------------------------------ tasks 3..13: terminations, reward terms
    if (!do_rew) continue;
    if (task >= 5 && !live) continue;
    if (task >= 11 && !has_obj) continue;
    const float* fmx = sm + L.fmax + e;  // [body][kEnvs]
    switch (task) {
      case 3:
      case 4: {  // terminations (terminations.py:10-23 + [IL] terms): task 3 takes the even terms, task 4 the odd ones; is_alive
        bool terminated = false, timed_out = false;
        if (live) {
#pragma unroll 1
          for (int t = task - 3; t < A.num_termination_terms; t += 2) {
            const LtTerminationTerm& tt = A.termination_terms[t];
            bool m = false;
            switch (tt.kind) {
              case LT_TK_TIME_OUT: m = reinterpret_cast<const long long*>(sm + L.eplen)[e] >= A.max_episode_length; break;
              case LT_TK_BAD_ORIENTATION: m = fabsf(acosf(-sm[L.grav + e * kS3 + 2])) > tt.p[0]; break;
              case LT_TK_ROOT_HEIGHT: m = sm[L.pos + e * kS3 + 2] < tt.p[0]; break;
              case LT_TK_ILLEGAL_CONTACT:
#pragma unroll 1
                for (int k = 0; k < tt.num_ids; ++k) m = m || fmx[tt.body_ids[k] * kEnvs] > tt.p[0];
                break;
              case LT_TK_OBJECT_BELOW_ROBOT: m = sm[L.opos + e * kS3 + 2] < sm[L.pos + e * kS3 + 2]; break;
              case LT_TK_BAD_ROLL: m = fabsf(asinf(sm[L.ograv + e * kS3 + 1])) > tt.p[0]; break;
            }
            if (A.term_masks) A.term_masks[(size_t)t * A.N + n] = m;
            if (tt.time_out) timed_out = timed_out || m; else terminated = terminated || m;
          }
          if (task == 4) s_tflag[e] = (unsigned char)((terminated ? 1 : 0) | (timed_out ? 2 : 0));
        }
        named_barrier(2, 64);
        if (task == 3 && live) {
          const unsigned char other = s_tflag[e];
          terminated = terminated || (other & 1);
          timed_out = timed_out || (other & 2);
          const bool done = terminated || timed_out;
          s_done[e] = done;
          s_tout[e] = timed_out;
          A.terminated[n] = terminated;
          A.time_outs[n] = timed_out;
          A.dones[n] = done;
          if (A.auto_reset && done) {  // history of a reset env is refilled by its next observation
            if (do_obs) s_fill[e] = 1;
            else if (A.obs_fill) A.obs_fill[n] = 1;
          }
          put(LT_RK_ALIVE, terminated ? 0.f : 1.f);
        }
        break;
      }
      case 5: {  // velocity tracking, base terms
        const Vec3 cmd = ld3(sm + L.cmd + e * kS3), vb = ld3(sm + L.linb + e * kS3), wb = ld3(sm + L.angb + e * kS3);
        const Vec3 grav = ld3(sm + L.grav + e * kS3);
        const float dx = cmd.x - vb.x, dy = cmd.y - vb.y;
        put(LT_RK_TRACK_LIN_VEL_XY, exp_neg_over(sqrtf(dx * dx + dy * dy), par(LT_RK_TRACK_LIN_VEL_XY, 0)));
        put(LT_RK_TRACK_ANG_VEL_Z, exp_neg_over(fabsf(cmd.z - wb.z), par(LT_RK_TRACK_ANG_VEL_Z, 0)));
        const float d = sm[L.pos + e * kS3 + 2] - par(LT_RK_BASE_HEIGHT, 0);
        put(LT_RK_BASE_HEIGHT, d * d);
        put(LT_RK_BASE_Z_VEL, vb.z * vb.z);
        put(LT_RK_BASE_RP_ANGLE, grav.x * grav.x + grav.y * grav.y);
        put(LT_RK_BASE_RP_VEL, fabsf(wb.x) + fabsf(wb.y));
        break;
      }
      case 6: {  // foot slipping / dragging
        const float* fpos = sm + L.fpos + e * kS12;
        const float* fvel = sm + L.fvel + e * kS12;
        const float slip_thr = par(LT_RK_FOOT_SLIP, 0), drag_h = par(LT_RK_FOOT_DRAG, 0), drag_v = par(LT_RK_FOOT_DRAG, 1);
        float slip = 0.f, drag = 0.f;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const float sp = sqrtf(fvel[3 * k] * fvel[3 * k] + fvel[3 * k + 1] * fvel[3 * k + 1]);
          slip += (fmx[A.feet_sensor_ids[k] * kEnvs] > slip_thr ? 1.f : 0.f) * sp;
          drag += (fpos[3 * k + 2] <= drag_h && sp > drag_v) ? 1.f : 0.f;
        }
        put(LT_RK_FOOT_SLIP, slip);
        put(LT_RK_FOOT_DRAG, drag);
        break;
      }
      case 7: {  // joint position limits / deviation
        const float* qq = sm + L.q + e * sJ; const float* q0 = sm + L.q0 + e * sJ; const float* lim = sm + L.lim + e * L.sLim;
        float s_lim = 0.f, s_dev = 0.f;
#pragma unroll 2
        for (int j = 0; j < J; ++j) {
          s_lim += -fminf(qq[j] - lim[2 * j], 0.f) + fmaxf(qq[j] - lim[2 * j + 1], 0.f);
          const float d = qq[j] - q0[j];
          s_dev += d * d;
        }
        put(LT_RK_JOINT_POS_LIMIT, s_lim);
        const Vec3 cmd = ld3(sm + L.cmd + e * kS3), vb = ld3(sm + L.linb + e * kS3);
        const float cmd_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z);
        const float dev = sqrtf(s_dev), bv = sqrtf(vb.x * vb.x + vb.y * vb.y);
        put(LT_RK_JOINT_POS, (cmd_norm > 0.f || bv > par(LT_RK_JOINT_POS, 1)) ? dev : par(LT_RK_JOINT_POS, 0) * dev);
        break;
      }
      case 8: {  // joint acceleration / velocity
        const float* qd = sm + L.qd + e * sJ; const float* qdd = sm + L.qdd + e * sJ;
        float s_acc = 0.f, s_vel = 0.f;
#pragma unroll 4
        for (int j = 0; j < J; ++j) {
          s_acc += qdd[j] * qdd[j];
          s_vel += qd[j] * qd[j];
        }
        put(LT_RK_JOINT_ACC, sqrtf(s_acc));
        put(LT_RK_JOINT_VEL, sqrtf(s_vel));
        break;
      }
      case 9: {  // joint torque, action rate
        const float* tau = sm + L.tau + e * sJ; const float* ac = sm + L.act + e * sJ; const float* pa = sm + L.pact + e * sJ;
        float s_tau = 0.f, s_rate = 0.f;
#pragma unroll 4
        for (int j = 0; j < J; ++j) {
          s_tau += tau[j] * tau[j];
          const float da = ac[j] - pa[j];
          s_rate += da * da;
        }
        put(LT_RK_JOINT_TORQUE, sqrtf(s_tau));
        put(LT_RK_ACTION_RATE, s_rate);
        break;
      }
      case 10: {  // thigh / calf collisions
        const float thr = par(LT_RK_THIGH_CALF_COLLISION, 0);
        float c = 0.f;
        if (s_slot[LT_RK_THIGH_CALF_COLLISION] >= 0) {
#pragma unroll 1
          for (int k = 0; k < A.num_thigh_calf; ++k) c += fmx[A.thigh_calf_sensor_ids[k] * kEnvs] > thr ? 1.f : 0.f;
        }
        put(LT_RK_THIGH_CALF_COLLISION, c);
        break;
      }
      case 11: {  // object position / linear velocity terms, dangerous state (rewards.py:469-503, 569-594)
        const Vec3 cmd = ld3(sm + L.cmd + e * kS3);
        const float cmd_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z);
        const float moving = cmd_norm > 0.f ? 1.f : 0.f;
        const Quat q = ld4(sm + L.quat + e * kS4);
        const Vec3 rel_pos_w = sub3(ld3(sm + L.opos + e * kS3), ld3(sm + L.pos + e * kS3));
        const Vec3 rel_vel = {s_of[3 * kEnvs + e], s_of[4 * kEnvs + e], s_of[5 * kEnvs + e]};
        {
          float v = sqrtf(rel_pos_w.x * rel_pos_w.x + rel_pos_w.y * rel_pos_w.y);
          if (par(LT_RK_OBJ_XY_POS, 0) != 0.f) v *= moving;
          put(LT_RK_OBJ_XY_POS, v);
        }
        put(LT_RK_OBJ_XY_VEL, rel_vel.x * rel_vel.x + rel_vel.y * rel_vel.y);
        put(LT_RK_OBJ_LOSE_CONTACT, (sm[L.oc_last + e] > 0.f && sm[L.oc_air + e] > 0.f) ? 1.f : 0.f);
        put(LT_RK_OBJ_Z_VEL, rel_vel.z * rel_vel.z);
        if (s_slot[LT_RK_OBJ_DANGER] >= 0) {
          const Vec3 rel_pos = {s_of[0 * kEnvs + e], s_of[1 * kEnvs + e], s_of[2 * kEnvs + e]};
          bool bad = fabsf(rel_pos.x) > par(LT_RK_OBJ_DANGER, 0);
          bad = bad || fabsf(rel_pos.y) > par(LT_RK_OBJ_DANGER, 1);
          bad = bad || rel_pos.z < par(LT_RK_OBJ_DANGER, 2);
          const float rp = par(LT_RK_OBJ_DANGER, 3), vmax = par(LT_RK_OBJ_DANGER, 4);
          if (rp >= 0.f) bad = bad || fabsf(acosf(-sm[L.ograv + e * kS3 + 2])) > rp * 3.14159265358979323846f / 180.f;
          if (vmax >= 0.f) bad = bad || sqrtf(rel_vel.x * rel_vel.x + rel_vel.y * rel_vel.y) > vmax;
          put(LT_RK_OBJ_DANGER, bad ? 1.f : 0.f);
        }
        break;
      }
      case 12: {  // object orientation / angular velocity terms (rewards.py:505-543)
        const Quat q = ld4(sm + L.quat + e * kS4);
        const Vec3 rel_ang = {s_of[6 * kEnvs + e], s_of[7 * kEnvs + e], s_of[8 * kEnvs + e]};
        const Vec3 g_obj = {s_of[9 * kEnvs + e], s_of[10 * kEnvs + e], s_of[11 * kEnvs + e]};
        put(LT_RK_OBJ_RP_ANGLE, g_obj.x * g_obj.x + g_obj.y * g_obj.y);
        put(LT_RK_OBJ_RP_VEL, fabsf(rel_ang.x) + fabsf(rel_ang.y));
        put(LT_RK_OBJ_ROLL_ANGLE, g_obj.y * g_obj.y);
        put(LT_RK_OBJ_ROLL_VEL, rel_ang.x * rel_ang.x);
        break;
      }
      case 13: {  // object yaw alignment (rewards.py:545-567)
        if (s_slot[LT_RK_OBJ_YAW] < 0) break;
        const Vec3 cmd = ld3(sm + L.cmd + e * kS3);
        const float moving = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z) > 0.f ? 1.f : 0.f;
        const Quat qr = yaw_quat(yaw_of(ld4(sm + L.quat + e * kS4))), qo = yaw_quat(yaw_of(ld4(sm + L.oquat + e * kS4)));
        float d = yaw_of(quat_mul(quat_inv(qr), qo));
        const float pi = 3.14159274101257324f;  // float32(torch.pi)
        if (d > pi) d -= 2.f * pi;
        if (d > 0.5f * pi) d -= pi;
        if (d <= -0.5f * pi) d += pi;
        float v = d * d;
        if (par(LT_RK_OBJ_YAW, 0) != 0.f) v *= moving;
        put(LT_RK_OBJ_YAW, v);
        break;
      }
      default: break;
    }
  }
#ifdef LT_MDP_TWICE
  if (pass == 0) PROF_STAMP(13);
  }
#endif
  PROF_STAMP(4);
  // ActionManager.reset(env_ids) of IsaacLab's step order (rewards -> reset -> observations; reference mdp/actions.py:46-52 and the
  // [IL] ActionTerm.reset behind it): the action history of an env that this launch resets is zeroed BEFORE its post-reset observation
  // is built.  processed_actions keeps raw * scale + offset of the pre-reset raw action (actions.py:49 runs before the base class
  // zeroes raw).  The new last_action values were computed by the task phase from the pre-reset row, so they are redone here for
  // the (few) done envs with raw = 0 and the same uniform of the same Philox word / u_obs column.
  if (A.act_reset_on_done && A.auto_reset && do_rew) {
    __syncthreads();  // s_done of every env of the block is final; nobody reads the staged action rows any more
#pragma unroll 1
    for (int i = tid; i < nvalid * J; i += kThreads) {
      const int ee = i / J, c = i - ee * J;
      if (!s_done[ee]) continue;
      const size_t g = (size_t)(e0 + ee) * J + c;
      const_cast<float*>(A.raw_actions)[g] = 0.f;
      const_cast<float*>(A.prev_raw_actions)[g] = 0.f;
      if (A.act_prev_prev_raw) A.act_prev_prev_raw[g] = 0.f;
      if (!do_obs) continue;
      int jb = 0;
#pragma unroll 1
      for (int t = 0; t < A.num_obs_terms; ++t) {
        const LtObsTerm& ot = A.obs_terms[t];
        if (ot.kind == LT_OK_LAST_ACTION) {
          const int j = jb + c;
          float noisy = 0.f;
          if (ot.noisy) {
            float u;
            if (A.u_obs) {
              u = __ldcs(A.u_obs + (size_t)(e0 + ee) * dps + j);
            } else {
              const uint4 r4 = philox4(A.seed, rng_offset, (uint32_t)(e0 + ee), (uint32_t)(j >> 2));
              const int w = j & 3;
              u = lt::Philox::u01(w == 0 ? r4.x : (w == 1 ? r4.y : (w == 2 ? r4.z : r4.w)));
            }
            noisy = (0.f + u * (ot.n_max - ot.n_min)) + ot.n_min;
          }
          nv[(0 * kEnvs + ee) * sNew + j] = noisy * ot.scale;
          nv[(1 * kEnvs + ee) * sNew + j] = 0.f * ot.scale;
        }
        jb += ot.dim;
      }
    }
  }
  if (do_obs) {
    __pipeline_wait_prior(0);
    if (bulk_hist) mbar_wait(&s_bar[1], 0);
  }
  PROF_STAMP(5);
  __syncthreads();
  PROF_STAMP(6);

  // ---------------------------------------------------------------------------- stage 2a: reward accumulation + outputs
  // The last warp owns the per-env reward (a sequential fp32 sum over the terms, the longest chain of this stage) and nothing
  // else; the other warps share the per-term outputs.
  const int a_warps = kWarps, b_first = 0;
  if (do_rew && warp < a_warps) {
    const float dt = A.step_dt;
    const int a_threads = (a_warps - 1) * 32;
    // [IL] RewardManager.compute: value = f * weight * dt ; episode_sums += value ; step_reward = value / dt
#pragma unroll 1
    for (int i = tid; i < T * kEnvs && warp < a_warps - 1; i += a_threads) {
      const int t = i >> 5, ee = i & 31, nn = e0 + ee;
      if (ee >= nvalid) continue;
      const LtRewardTerm& rt = A.reward_terms[t];
      if (rt.weight == 0.f) {  // skipped terms: step_reward column is zero, sums untouched
        if (A.step_reward) A.step_reward[(size_t)nn * T + t] = 0.f;
        continue;
      }
      const float raw = s_raw[rt.kind * kEnvs + ee];
      const float value = (raw * rt.weight) * dt;
      if (A.term_raw) A.term_raw[(size_t)t * A.N + nn] = raw;
      if (A.step_reward) A.step_reward[(size_t)nn * T + t] = value / dt;
      if (A.episode_sums) {
        const float total = sm[L.esum + i] + value;
        const bool rst = A.auto_reset && s_done[ee];
        if (rst && A.episode_log_sums) atomicAdd(A.episode_log_sums + t, total);
        A.episode_sums[(size_t)t * A.N + nn] = rst ? 0.f : total;
      }
    }
    PROF_STAMP(19);
    if (warp == a_warps - 1 && live) {  // reward_buf: sequential fp32 sum in term order
      // Branch-free body: kind / weight / raw value are loaded unconditionally (a skipped term reads a stale slot and is
      // replaced by 0 with a select), so the loads of a whole unrolled group are in flight together and only the additions,
      // which must keep the term order, form a chain.  With a branch per term every term was its own serial
      // constant-load -> branch -> constant-load -> shared-load -> multiply chain (~200 cycles each, 2.3 us in total).
      float reward = 0.f;
      int t = 0;
#pragma unroll 1
      for (; t + 8 <= T; t += 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float w = A.reward_terms[t + k].weight;
          const float raw = s_raw[A.reward_terms[t + k].kind * kEnvs + e];
          v[k] = w != 0.f ? (raw * w) * dt : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) reward = reward + v[k];
      }
#pragma unroll 1
      for (; t < T; ++t) {
        const float w = A.reward_terms[t].weight;
        const float raw = s_raw[A.reward_terms[t].kind * kEnvs + e];
        reward = reward + (w != 0.f ? (raw * w) * dt : 0.f);
      }
      A.reward[n] = reward;
      if (A.store_rewards) {  // K3 inside K1: time-out bootstrap (ppo.py:162-165) + the scalar rollout store, lt_store_step bit for bit
        float r = reward;
        if (A.store_values) r = __fadd_rn(r, __fmul_rn(A.store_gamma, __fmul_rn(A.store_values[n], s_tout[e] ? 1.f : 0.f)));
        A.store_rewards[n] = r;
        if (A.store_dones) A.store_dones[n] = s_done[e] ? 1 : 0;
      }
      if (A.auto_reset && s_done[e] && A.episode_log_sums) atomicAdd(A.episode_log_sums + T, 1.0f);
    }
    PROF_STAMP(20);
    // gait state write back; zeroed when the env is being reset (the manager's reset(env_ids) follows, rewards.py:107-114)
    {
      const LtGaitState& G = A.gait_state;
#pragma unroll 1
      for (int i = tid; i < nvalid * kGaitFloats && warp < a_warps - 1; i += a_threads) {
        const int ee = i / kGaitFloats, k = i - ee * kGaitFloats, nn = e0 + ee;
        const float v = (A.auto_reset && s_done[ee]) ? 0.f : sm[L.gait_out + ee * kGaitStride + k];
        if (k < 4) G.last_step_current_air_time[nn * 4 + k] = v;
        else if (k < 8) G.last_step_current_contact_time[nn * 4 + k - 4] = v;
        else if (k < 12) G.valid_last_air_time[nn * 4 + k - 8] = v;
        else if (k < 15) G.last_velocity_cmd[nn * 3 + k - 12] = v;
        else if (k == 15) G.step_from_changing_cmd[nn] = v;
        else if (k < 20) G.swinging_in_zero_cmd[nn * 4 + k - 16] = v != 0.f;
        else G.valid_previous_contact[nn * 4 + k - 20] = v != 0.f;
      }
    }
  }

  PROF_STAMP(7);
  // --------------------------------------------------------------------------------- stage 2b: observation history shift
  // one warp per env row: rows are 8-byte aligned (D is even), so every lane moves float2 pairs
  if (do_obs && warp >= b_first) {
#pragma unroll 1
    for (int ee = warp - b_first; ee < nvalid; ee += kWarps - b_first) {
      const bool fill = s_fill[ee] != 0;
#pragma unroll 1
      for (int grp = 0; grp < 2; ++grp) {
        float* out = grp ? A.critic_obs_out : A.policy_obs_out;
        if (!out) continue;
        const bool refill = fill || (grp ? A.critic_obs_in : A.policy_obs_in) == nullptr;
        const float* hist = sm + L.hist + grp * L.hist_stride + ee * D;
        const float* nve = nv + (grp * kEnvs + ee) * sNew;
        float* dst = out + (size_t)(e0 + ee) * D;
        if ((D & 1) == 0 && (((uintptr_t)dst) & 7) == 0) {
#pragma unroll 4
          for (int p = lane; p < (D >> 1); p += 32) {
            const uint2 m = *reinterpret_cast<const uint2*>(s_map + 2 * p);
            float2 v;
            v.x = (refill || (m.x >> 16) == 0xffff) ? nve[m.x & 0xffff] : hist[m.x >> 16];
            v.y = (refill || (m.y >> 16) == 0xffff) ? nve[m.y & 0xffff] : hist[m.y >> 16];
            __stcs(reinterpret_cast<float2*>(dst) + p, v);
          }
        } else {
#pragma unroll 1
          for (int k = lane; k < D; k += 32) {
            const unsigned m = (unsigned)s_map[k];
            __stcs(dst + k, (refill || (m >> 16) == 0xffff) ? nve[m & 0xffff] : hist[m >> 16]);
          }
        }
      }
      if (A.obs_fill && lane == 0) A.obs_fill[e0 + ee] = 0;
    }
  }
  PROF_STAMP(15);
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

// ints in the table block: s_map[D rounded up to 4] + s_jinfo[dps] + s_slot[LT_RK_COUNT] + s_kpar[LT_RK_COUNT][6] (floats), rounded
// up to a multiple of 4
int tables_len(int D, int dps) { return ((((D + 3) & ~3) + dps + LT_RK_COUNT * 7) + 3) & ~3; }

// validates the observation term table; *dps = new values per step and group
int obs_dims(const LtMdpArgs* a, int* dps_out, bool pointers_bound) {
  const bool has_obj = !pointers_bound || a->obj_root_pos_w != nullptr;  // table builders run before the tensors are bound
  if (a->num_obs_terms <= 0 || a->num_obs_terms > LT_MAX_OBS_TERMS || a->history_length <= 0) return LT_ERR_INVALID_ARG;
  int dps = 0;
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
  if (dps > kMaxNew || dps * a->history_length > kMaxObsDim || dps > 255) return LT_ERR_UNSUPPORTED;
  *dps_out = dps;
  return LT_OK;
}

}  // namespace

// Launch-constant lookup tables of the fused MDP kernel (observation column map, per-value term info, reward kind -> slot),
// computed once on the host; upload the ints to the device and pass them as LtMdpArgs.tables.  Without them every block of
// every launch rebuilds the tables in shared memory.
extern "C" int lt_mdp_tables_len(const LtMdpArgs* a) {
  if (!a) return -1;
  int dps = 0;
  if (a->num_obs_terms > 0 && obs_dims(a, &dps, false) != LT_OK) return -1;
  return tables_len(dps * a->history_length, dps);
}

extern "C" int lt_mdp_build_tables(const LtMdpArgs* a, int32_t* out, int len) {
  if (!a || !out) return LT_ERR_INVALID_ARG;
  int dps = 0;
  if (a->num_obs_terms > 0) {
    const int rc = obs_dims(a, &dps, false);
    if (rc != LT_OK) return rc;
  }
  const int D = dps * a->history_length;
  if (len < tables_len(D, dps)) return LT_ERR_INVALID_ARG;
  memset(out, 0, sizeof(int32_t) * (size_t)len);
  int32_t* map = out;
  int32_t* jinfo = out + ((D + 3) & ~3);
  int32_t* slot = jinfo + dps;
  int col = 0, jbase = 0;
  for (int t = 0; t < a->num_obs_terms; ++t) {  // flattened [term][history][dim] row
    const int d = a->obs_terms[t].dim;
    for (int h = 0; h < a->history_length; ++h)
      for (int i = 0; i < d; ++i) {
        const int k = col + h * d + i;
        map[k] = (jbase + i) | ((h == a->history_length - 1 ? 0xffff : k + d) << 16);
      }
    for (int i = 0; i < d; ++i) jinfo[jbase + i] = t | (i << 8);
    col += d * a->history_length;
    jbase += d;
  }
  for (int kind = 0; kind < LT_RK_COUNT; ++kind) {
    int sl = -1;
    for (int i = 0; i < a->num_reward_terms && i < LT_MAX_REWARD_TERMS; ++i)
      if (a->reward_terms[i].kind == kind && a->reward_terms[i].weight != 0.f) sl = i;
    slot[kind] = sl;
    for (int k = 0; k < 6; ++k) {
      const float v = sl >= 0 ? a->reward_terms[sl].p[k] : 0.f;
      memcpy(slot + LT_RK_COUNT + kind * 6 + k, &v, sizeof(float));
    }
  }
  return LT_OK;
}

extern "C" int lt_mdp_step(const LtMdpArgs* a, void* stream) {
  if (!a || a->N <= 0 || !(a->phases & (LT_PHASE_REWARDS | LT_PHASE_OBS))) return LT_ERR_INVALID_ARG;
  const bool do_rew = a->phases & LT_PHASE_REWARDS, do_obs = a->phases & LT_PHASE_OBS;
  const bool has_obj = a->obj_root_pos_w != nullptr;
  if (a->J <= 0 || a->J > kMaxJ) return LT_ERR_INVALID_ARG;
  if (a->act_new && !do_rew) return LT_ERR_INVALID_ARG;  // the fused action term belongs to the reward pass (the previous raw actions are staged there)
  if (a->store_rewards && !do_rew) return LT_ERR_INVALID_ARG;  // the fused rollout store writes what the reward pass computes
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
    if (a->num_obs_terms <= 0 || !a->default_joint_vel) return LT_ERR_INVALID_ARG;
    if (!a->policy_obs_out && !a->critic_obs_out) return LT_ERR_INVALID_ARG;
  }
  if (a->num_obs_terms > 0) {  // the table block is laid out for the full configuration, whichever phases this launch runs
    const int rc = obs_dims(a, &dps, true);
    if (rc != LT_OK) return rc;
    L.D = dps * a->history_length;
    L.dps = dps;
  }
  const int J = a->J, S = a->num_sensor_bodies > 0 ? a->num_sensor_bodies : 1, H = a->force_history > 0 ? a->force_history : 1;
  const int T = do_rew ? a->num_reward_terms : 0;
  L.sJ = J; L.sLim = 2 * J; L.sF = H * S * 3;
  L.sNew = (dps > 0 ? dps : 1) | 1;
  int off = 0;
  auto take = [&](int per_env) { const int o = off; off += kEnvs * per_env; return o; };
  L.cmd = take(kS3); L.pos = take(kS3); L.linb = take(kS3); L.angb = take(kS3); L.grav = take(kS3);
  L.q = take(L.sJ); L.qd = take(L.sJ); L.qdd = take(L.sJ); L.tau = take(L.sJ); L.q0 = take(L.sJ); L.qd0 = take(L.sJ); L.lim = take(L.sLim);
  L.act = take(L.sJ); L.pact = take(L.sJ);
  L.force = take(do_rew ? L.sF : 0);
  L.air = take(kP4); L.con = take(kP4); L.lair = take(kP4); L.fpos = take(kS12); L.fvel = take(kS12);
  L.quat = take(kS4); L.linw = take(kS3); L.angw = take(kS3); L.opos = take(kS3); L.oquat = take(kS4); L.olin = take(kS3); L.oang = take(kS3);
  L.ograv = take(kS3); L.oc_last = take(1); L.oc_cur = take(1); L.oc_air = take(1);
  L.g_lsa = take(kS4); L.g_lsc = take(kS4); L.g_vla = take(kS4); L.g_cmd = take(kS3); L.g_steps = take(1); L.g_sz = take(kP4); L.g_vpc = take(kP4);
  L.eplen = take(2);  // int64 per env
  L.gait_out = take(kGaitStride);
  L.esum = take(T); L.raw = take(LT_RK_COUNT);  // [T][kEnvs], [kinds][kEnvs]
  L.gtmp = take(4);
  L.fmax = take(do_rew ? S : 0);  // [S][kEnvs]
  L.oframe = take(has_obj ? 16 : 0);  // [16][kEnvs]: object position / velocity / angular velocity / gravity / quaternion in the robot frame
  L.newobs = take(2 * L.sNew);
  L.hist_stride = ((kEnvs * L.D + 3) & ~3) + 4;  // +4: the shifted read of the last element may touch one slot past the block
  L.hist = off; off += do_obs ? 2 * L.hist_stride : 0;
  L.tables_len = tables_len(L.D, dps);
  L.map = off; off += L.tables_len;
  L.total = off;
  const size_t smem = (size_t)L.total * sizeof(float);
  if (smem > 220 * 1024) return LT_ERR_UNSUPPORTED;

  StageTable ST;
  memset(&ST, 0, sizeof(ST));
  bool overflow = false, aligned = true;
  unsigned tx = 0;
  auto add = [&](const float* src, int row, int dst_off) {
    if (!src || row <= 0) return;
    if (ST.n >= kMaxStage) { overflow = true; return; }
    const int t = ST.n++;
    ST.src[t] = src; ST.row[t] = row; ST.off[t] = dst_off;
    aligned = aligned && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    tx += (unsigned)(kEnvs * 4) * (unsigned)row;
  };
  if (do_rew) add(a->net_forces_w_history, H * S * 3, L.force);  // biggest first
  add(a->joint_pos, J, L.q); add(a->joint_vel, J, L.qd); add(a->default_joint_pos, J, L.q0); add(a->raw_actions, J, L.act);
  if (do_obs) add(a->default_joint_vel, J, L.qd0);
  if (do_rew) {
    add(a->soft_joint_pos_limits, 2 * J, L.lim); add(a->joint_acc, J, L.qdd); add(a->applied_torque, J, L.tau);
    add(a->prev_raw_actions, J, L.pact);
  }
  add(a->command, 3, L.cmd); add(a->root_pos_w, 3, L.pos); add(a->root_ang_vel_b, 3, L.angb); add(a->projected_gravity_b, 3, L.grav);
  if (do_rew) {
    add(a->root_lin_vel_b, 3, L.linb);
    const LtGaitState& gs = a->gait_state;
    add(gs.last_step_current_air_time, 4, L.g_lsa); add(gs.last_step_current_contact_time, 4, L.g_lsc); add(gs.valid_last_air_time, 4, L.g_vla);
    add(gs.last_velocity_cmd, 3, L.g_cmd); add(gs.step_from_changing_cmd, 1, L.g_steps);
  }
  if (has_obj) {
    if (!a->obj_last_contact_time || !a->obj_current_contact_time || !a->obj_current_air_time) return LT_ERR_INVALID_ARG;
    add(a->root_quat_w, 4, L.quat); add(a->obj_root_quat_w, 4, L.oquat);
    add(a->obj_last_contact_time, 1, L.oc_last); add(a->obj_current_contact_time, 1, L.oc_cur); add(a->obj_current_air_time, 1, L.oc_air);
    add(a->root_lin_vel_w, 3, L.linw); add(a->root_ang_vel_w, 3, L.angw);
    add(a->obj_root_pos_w, 3, L.opos); add(a->obj_root_lin_vel_w, 3, L.olin);
    add(a->obj_root_ang_vel_w, 3, L.oang); add(a->obj_projected_gravity_b, 3, L.ograv);
  }
  if (overflow) return LT_ERR_UNSUPPORTED;
  if (a->tables) {
    aligned = aligned && (reinterpret_cast<uintptr_t>(a->tables) & 15) == 0;
    tx += (unsigned)L.tables_len * 4u;
  }
  L.bulk_ok = aligned ? 1 : 0;
  L.tx_state = tx;
  L.hist_bulk_ok = do_obs && (!a->policy_obs_in || (reinterpret_cast<uintptr_t>(a->policy_obs_in) & 15) == 0) &&
                   (!a->critic_obs_in || (reinterpret_cast<uintptr_t>(a->critic_obs_in) & 15) == 0) && ((size_t)L.D * kEnvs * 4) % 16 == 0;

  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mdp_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
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

#ifdef LT_MDP_PROF
extern "C" int lt_debug_mdp_prof(void* buf) { return lt::check(cudaMemcpyToSymbol(g_mdp_prof, &buf, sizeof(buf))); }
#endif

extern "C" int lt_mdp_reset(const LtGaitState* g, float* episode_sums, int num_reward_terms, const uint8_t* mask, int N, void* stream) {
  if (!g || !mask || N <= 0) return LT_ERR_INVALID_ARG;
  mdp_reset_kernel<<<(unsigned)lt::ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(*g, episode_sums, num_reward_terms, mask, N);
  return lt::check_launch();
}
