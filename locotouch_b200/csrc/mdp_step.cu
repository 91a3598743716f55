// K1: fused MDP step -- terminations -> rewards (incl. the stateful adaptive symmetric gait reward) -> [auto reset]
// -> policy / critic observations with history shift and noise.
// Replaces the per-term PyTorch code of reference locotouch/mdp/rewards.py:15-604, terminations.py:10-23,
// observations.py:38-91 and the IsaacLab manager loops around it (several thousand ATen launches per env step,
// SURVEY.md 2.1) with one launch that reads exactly the IsaacLab tensors the reference terms read.
//
// Block = 128 threads = kEnvs consecutive envs.
//   stage 0  all threads copy the block's rows of every input tensor into shared memory with coalesced accesses
//            (the rows of kEnvs consecutive envs are one contiguous span in each [N, row] tensor).
//   stage 1  all threads: per-(env, body) contact force maxima; per-(env, element) new observation values (noise, scale).
//   stage 2  warp 0, one lane per env: termination terms.
//   stage 3  warp 0, one lane per env: reward terms in manager order, gait state update, episode sums, auto reset;
//            warps 1..3 at the same time: observation history shift, all loads issued before the first store.
// Floating point follows the reference's fp32 expression order; masks are bit-exact by construction.
#include <math_constants.h>
#include <string.h>

#include "lt_common.cuh"

namespace {

constexpr int kEnvs = 4;        // envs per block
constexpr int kThreads = 128;
constexpr int kObsThreads = kThreads - 32;
constexpr int kMaxObsRegs = 32;  // history values per obs thread: 2 groups * kEnvs * D / 96 <= 32  ->  D <= 384
constexpr int kMaxObsDim = 384;
constexpr int kMaxNew = 64;      // new observation values per env and group
constexpr int kMaxJ = 16;
constexpr int kMaxSensorBodies = 32;

// shared-memory layout (float offsets, per-env row stride = row length), filled in on the host
struct Layout {
  int cmd, pos, linb, angb, grav, q, qd, qdd, tau, q0, qd0, lim, act, pact, force, air, con, lair, fpos, fvel;
  int quat, linw, angw, opos, oquat, olin, oang, ograv, octime, fmax, gait, newobs, map, misc, total;
  int D;    // observation dim per group
  int dps;  // new values per step per group
};

struct Vec3 { float x, y, z; };
struct Quat { float w, x, y, z; };

__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// [IL] isaaclab.utils.math.quat_apply / quat_apply_inverse
__device__ __forceinline__ Vec3 quat_apply(Quat q, Vec3 v, float sign) {
  const Vec3 u = {q.x, q.y, q.z};
  Vec3 t = cross(u, v);
  t = {t.x * 2.f, t.y * 2.f, t.z * 2.f};
  const Vec3 c = cross(u, t);
  return {v.x + sign * q.w * t.x + c.x, v.y + sign * q.w * t.y + c.y, v.z + sign * q.w * t.z + c.z};
}
__device__ __forceinline__ Vec3 rot(Quat q, Vec3 v) { return quat_apply(q, v, 1.f); }
__device__ __forceinline__ Vec3 rot_inv(Quat q, Vec3 v) { return quat_apply(q, v, -1.f); }
// [IL] quat_mul, 8-multiply factored form
__device__ __forceinline__ Quat quat_mul(Quat a, Quat b) {
  const float ww = (a.z + a.x) * (b.x + b.y);
  const float yy = (a.w - a.y) * (b.w + b.z);
  const float zz = (a.w + a.y) * (b.w - b.z);
  const float xx = ww + yy + zz;
  const float qq = 0.5f * (xx + (a.z - a.x) * (b.x - b.y));
  return {qq - ww + (a.z - a.y) * (b.y - b.z), qq - xx + (a.x + a.w) * (b.x + b.w), qq - yy + (a.w - a.x) * (b.y + b.z),
          qq - zz + (a.z + a.y) * (b.w - b.x)};
}
// [IL] quat_inv: conj(q) / max(|q|^2, 1e-9)
__device__ __forceinline__ Quat quat_inv(Quat q) {
  const float n = fmaxf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z, 1e-9f);
  return {q.w / n, -q.x / n, -q.y / n, -q.z / n};
}
__device__ __forceinline__ Quat quat_from_euler(float roll, float pitch, float yaw) {
  float sy, cy, sr, cr, sp, cp;
  sincosf(yaw * 0.5f, &sy, &cy);
  sincosf(roll * 0.5f, &sr, &cr);
  sincosf(pitch * 0.5f, &sp, &cp);
  return {cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp};
}
__device__ __forceinline__ float yaw_of(Quat q) { return atan2f(2.0f * (q.w * q.z + q.x * q.y), 1.f - 2.f * (q.y * q.y + q.z * q.z)); }
__device__ __forceinline__ Quat yaw_quat(float yaw) {
  float s, c;
  sincosf(yaw * 0.5f, &s, &c);
  return {c, 0.f, 0.f, s};
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ Vec3 ld3(const float* p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ Quat ld4(const float* p) { return {p[0], p[1], p[2], p[3]}; }
__device__ __forceinline__ Vec3 sub3(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }

__device__ __forceinline__ void named_barrier_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ---------------------------------------------------------------------------------------------- gait (rewards.py:60-392)
struct GaitRegs {
  float lsa[4], lsc[4], vla[4], last_cmd[3], steps;
  bool sz[4], vpc[4];
};

// rewards.py:158-200 in statement order
__device__ void gait_update(GaitRegs& g, const float a[4], const float c[4], const float la[4], Vec3 cmd, bool nz, bool any_nz, float th) {
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

// rewards.py:243-346.  pair = 0: feet (0,1) target, (2,3) other; pair = 1: the reverse.
__device__ float gait_swing_bonus(const LtGaitParams& gp, const GaitRegs& g, const float a[4], int pair) {
  const float th = gp.judge_time_threshold;
  const int t0 = pair ? 2 : 0, o0 = pair ? 0 : 2;
  const float a0 = a[t0], a1 = a[t0 + 1];
  const float m = (a0 + a1) / 2.f;
  const bool both_air = a0 > th && a1 > th;
  const float vt0 = g.vla[t0], vt1 = g.vla[t0 + 1], vo0 = g.vla[o0], vo1 = g.vla[o0 + 1];
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

// rewards.py:218-241
__device__ float gait_sync(const LtGaitParams& gp, const GaitRegs& g, const float a[4], const float c[4], int pair, float score) {
  const float th = gp.judge_time_threshold;
  const int f0 = pair ? 2 : 0, f1 = f0 + 1;
  const bool both_air = a[f0] > th && a[f0] < gp.air_time_gait_bound && a[f1] > th && a[f1] < gp.air_time_gait_bound;
  const bool c0 = c[f0] > th && c[f0] < gp.contact_time_gait_bound;
  const bool c1 = c[f1] > th && c[f1] < gp.contact_time_gait_bound;
  const bool both_contact = c0 && c1;
  if (gp.encourage_symmetricity) {
    float bonus = gait_swing_bonus(gp, g, a, pair);
    const float scale = 1.f - gp.task_performance_ratio + gp.task_performance_ratio * score;
    if (bonus > 0.f) bonus *= scale;
    bonus += 1.f;
    return both_air ? bonus : (both_contact ? 1.f : 0.f);
  }
  return (both_air || both_contact) ? 1.f : 0.f;
}

// rewards.py:348-363
__device__ float gait_async(const LtGaitParams& gp, const float a[4], const float c[4], int f0, int f1) {
  const float th = gp.judge_time_threshold, tha = gp.async_judge_time_threshold;
  const bool both = c[f0] > th && c[f0] <= tha && c[f1] > th && c[f1] <= tha;
  const bool a0 = a[f0] > th && a[f0] < gp.air_time_gait_bound, a1 = a[f1] > th && a[f1] < gp.air_time_gait_bound;
  const bool c0 = c[f0] > th && c[f0] < gp.contact_time_gait_bound, c1 = c[f1] > th && c[f1] < gp.contact_time_gait_bound;
  return (both || (a0 && c1) || (c0 && a1)) ? 1.f : 0.f;
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

__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, int e0, int nvalid, int row) {
  if (!src) return;
  const float* s = src + (size_t)e0 * row;
  const int count = nvalid * row;
  for (int i = threadIdx.x; i < count; i += kThreads) dst[i] = __ldcs(s + i);
}

__global__ void __launch_bounds__(kThreads) mdp_step_kernel(const LtMdpArgs A, const Layout L) {
  extern __shared__ float sm[];
  __shared__ unsigned char s_done[kEnvs], s_term[kEnvs], s_fill[kEnvs];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e0 = blockIdx.x * kEnvs;
  const int nvalid = min(kEnvs, A.N - e0);
  const bool do_rew = A.phases & LT_PHASE_REWARDS, do_obs = A.phases & LT_PHASE_OBS;
  const bool has_obj = A.obj_root_pos_w != nullptr;
  const int J = A.J, S = A.num_sensor_bodies, H = A.force_history;
  const int step = (int)A.offset;

  // ------------------------------------------------------------------------------------------------ stage 0: loads
  stage_rows(sm + L.cmd, A.command, e0, nvalid, 3);
  stage_rows(sm + L.angb, A.root_ang_vel_b, e0, nvalid, 3);
  stage_rows(sm + L.grav, A.projected_gravity_b, e0, nvalid, 3);
  stage_rows(sm + L.q, A.joint_pos, e0, nvalid, J);
  stage_rows(sm + L.qd, A.joint_vel, e0, nvalid, J);
  stage_rows(sm + L.q0, A.default_joint_pos, e0, nvalid, J);
  stage_rows(sm + L.act, A.raw_actions, e0, nvalid, J);
  stage_rows(sm + L.pos, A.root_pos_w, e0, nvalid, 3);
  if (do_obs) stage_rows(sm + L.qd0, A.default_joint_vel, e0, nvalid, J);
  if (do_rew) {
    stage_rows(sm + L.linb, A.root_lin_vel_b, e0, nvalid, 3);
    stage_rows(sm + L.qdd, A.joint_acc, e0, nvalid, J);
    stage_rows(sm + L.tau, A.applied_torque, e0, nvalid, J);
    stage_rows(sm + L.lim, A.soft_joint_pos_limits, e0, nvalid, 2 * J);
    stage_rows(sm + L.pact, A.prev_raw_actions, e0, nvalid, J);
    stage_rows(sm + L.force, A.net_forces_w_history, e0, nvalid, H * S * 3);
    // feet rows of the contact timers / body states only (ids are arbitrary)
    for (int i = tid; i < nvalid * 4; i += kThreads) {
      const int e = i >> 2, k = i & 3;
      const size_t r = (size_t)(e0 + e) * S + A.gait.feet_ids[k];
      sm[L.air + i] = __ldcs(A.current_air_time + r);
      sm[L.con + i] = __ldcs(A.current_contact_time + r);
      sm[L.lair + i] = __ldcs(A.last_air_time + r);
    }
    for (int i = tid; i < nvalid * 12; i += kThreads) {
      const int e = i / 12, k = (i % 12) / 3, c = i % 3;
      const size_t r = ((size_t)(e0 + e) * A.num_bodies + A.feet_body_ids[k]) * 3 + c;
      sm[L.fpos + i] = __ldcs(A.body_pos_w + r);
      sm[L.fvel + i] = __ldcs(A.body_lin_vel_w + r);
    }
    // gait state: 4+4+4+3+1 floats, 4+4 bytes per env
    const LtGaitState& G = A.gait_state;
    float* gs = sm + L.gait;
    for (int i = tid; i < nvalid * 24; i += kThreads) {
      const int e = i / 24, k = i % 24;
      const int n = e0 + e;
      float v;
      if (k < 4) v = G.last_step_current_air_time[n * 4 + k];
      else if (k < 8) v = G.last_step_current_contact_time[n * 4 + k - 4];
      else if (k < 12) v = G.valid_last_air_time[n * 4 + k - 8];
      else if (k < 15) v = G.last_velocity_cmd[n * 3 + k - 12];
      else if (k == 15) v = G.step_from_changing_cmd[n];
      else if (k < 20) v = (float)G.swinging_in_zero_cmd[n * 4 + k - 16];
      else v = (float)G.valid_previous_contact[n * 4 + k - 20];
      gs[i] = v;
    }
  }
  if (has_obj) {
    stage_rows(sm + L.quat, A.root_quat_w, e0, nvalid, 4);
    stage_rows(sm + L.linw, A.root_lin_vel_w, e0, nvalid, 3);
    stage_rows(sm + L.angw, A.root_ang_vel_w, e0, nvalid, 3);
    stage_rows(sm + L.opos, A.obj_root_pos_w, e0, nvalid, 3);
    stage_rows(sm + L.oquat, A.obj_root_quat_w, e0, nvalid, 4);
    stage_rows(sm + L.olin, A.obj_root_lin_vel_w, e0, nvalid, 3);
    stage_rows(sm + L.oang, A.obj_root_ang_vel_w, e0, nvalid, 3);
    stage_rows(sm + L.ograv, A.obj_projected_gravity_b, e0, nvalid, 3);
    for (int i = tid; i < nvalid * 3; i += kThreads) {
      const int e = i / 3, k = i % 3;
      const float* src = k == 0 ? A.obj_last_contact_time : (k == 1 ? A.obj_current_contact_time : A.obj_current_air_time);
      sm[L.octime + i] = src ? __ldcs(src + e0 + e) : 0.f;
    }
  }
  // observation map: for every position k of the flattened [term][history][dim] row, the index j of the per-step value
  // feeding that column (low 8 bits), the term width d (bits 8..15) and whether k is the newest history slot (bit 16)
  int* s_map = reinterpret_cast<int*>(sm + L.map);
  if (do_obs) {
    for (int k = tid; k < L.D; k += kThreads) {
      int col = 0, jbase = 0, entry = 0;
      for (int t = 0; t < A.num_obs_terms; ++t) {
        const int d = A.obs_terms[t].dim, span = d * A.history_length;
        if (k < col + span) {
          const int h = (k - col) / d, i = (k - col) % d;
          entry = (jbase + i) | (d << 8) | ((h == A.history_length - 1) << 16);
          break;
        }
        col += span;
        jbase += d;
      }
      s_map[k] = entry;
    }
    if (tid < nvalid) s_fill[tid] = A.obs_fill ? A.obs_fill[e0 + tid] : 0;
  }
  __syncthreads();

  // ------------------------------------------------------------------------- stage 1: contact maxima, new observations
  if (do_rew) {
    for (int i = tid; i < nvalid * S; i += kThreads) {
      const int e = i / S, b = i % S;
      const float* f = sm + L.force + e * (H * S * 3);
      float m = 0.f;
      for (int h = 0; h < H; ++h) {
        const float* p = f + (h * S + b) * 3;
        m = fmaxf(m, sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));  // torch.max(norm(F, dim=-1), dim=1)
      }
      sm[L.fmax + i] = m;
    }
  }
  if (do_obs) {
    float* nv = sm + L.newobs;  // [2 groups][kEnvs][dps]
    const int dps = L.dps;
    for (int i = tid; i < nvalid * dps; i += kThreads) {
      const int e = i / dps, j = i % dps, n = e0 + e;
      // locate the term
      int t = 0, jb = 0;
      while (j >= jb + A.obs_terms[t].dim) { jb += A.obs_terms[t].dim; ++t; }
      const LtObsTerm ot = A.obs_terms[t];
      if (ot.kind == LT_OK_OBJECT_STATE) continue;  // handled below
      const int c = j - jb;
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
        float u;
        if (A.u_obs) u = __ldcs(A.u_obs + (size_t)n * dps + j);
        else {
          const uint4 r = lt::Philox::gen(A.seed, A.offset, (uint32_t)n, (uint32_t)(j >> 2));
          const uint32_t w = (j & 3) == 0 ? r.x : ((j & 3) == 1 ? r.y : ((j & 3) == 2 ? r.z : r.w));
          u = lt::Philox::u01(w);
        }
        noisy = __fadd_rn(__fadd_rn(raw, __fmul_rn(u, ot.n_max - ot.n_min)), ot.n_min);  // [IL] data + rand*(max-min) + min
      }
      nv[(0 * kEnvs + e) * dps + j] = __fmul_rn(noisy, ot.scale);
      nv[(1 * kEnvs + e) * dps + j] = __fmul_rn(raw, ot.scale);
    }
    // object_state_in_robot_frame (observations.py:38-91): one thread per (env, group)
    if (has_obj) {
      int jb = 0, tobj = -1;
      for (int t = 0; t < A.num_obs_terms; ++t) {
        if (A.obs_terms[t].kind == LT_OK_OBJECT_STATE) { tobj = t; break; }
        jb += A.obs_terms[t].dim;
      }
      if (tobj >= 0 && tid < nvalid * 2) {
        const int e = tid >> 1, grp = tid & 1, n = e0 + e;
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
          float u[13], ue[3];
          if (A.u_obs) {
#pragma unroll
            for (int k = 0; k < 13; ++k) u[k] = __ldcs(A.u_obs + (size_t)n * dps + jb + k);
          } else {
#pragma unroll
            for (int k = 0; k < 13; ++k) {
              const int j = jb + k;
              const uint4 r = lt::Philox::gen(A.seed, A.offset, (uint32_t)n, (uint32_t)(j >> 2));
              u[k] = lt::Philox::u01((j & 3) == 0 ? r.x : ((j & 3) == 1 ? r.y : ((j & 3) == 2 ? r.z : r.w)));
            }
          }
          if (A.u_obj_euler) { ue[0] = A.u_obj_euler[n * 3]; ue[1] = A.u_obj_euler[n * 3 + 1]; ue[2] = A.u_obj_euler[n * 3 + 2]; }
          else {
            const uint4 r = lt::Philox::gen(A.seed, A.offset, (uint32_t)n, 0x40000000u);
            ue[0] = lt::Philox::u01(r.x); ue[1] = lt::Philox::u01(r.y); ue[2] = lt::Philox::u01(r.z);
          }
          float de[3];
#pragma unroll
          for (int k = 0; k < 13; ++k) {
            const bool is_q = k >= 6 && k < 10;
            const float lo = is_q ? 0.f : A.os_n_min[k], hi = is_q ? 0.f : A.os_n_max[k];
            const float add = __fadd_rn(__fmul_rn(u[k], hi - lo), lo);  // observations.py:77
            st[k] = __fadd_rn(st[k], add);
            cst[k] = __fadd_rn(cst[k], add);                            // observations.py:82 (same draw)
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) de[k] = __fadd_rn(__fmul_rn(ue[k], A.os_n_max[6 + k] - A.os_n_min[6 + k]), A.os_n_min[6 + k]);
          const Quat nq = quat_from_euler(de[0], de[1], de[2]);  // observations.py:78-79
          const Quat a = quat_mul({st[6], st[7], st[8], st[9]}, nq);
          st[6] = a.w; st[7] = a.x; st[8] = a.y; st[9] = a.z;
          const Quat b = quat_mul({cst[6], cst[7], cst[8], cst[9]}, nq);
          cst[6] = b.w; cst[7] = b.x; cst[8] = b.y; cst[9] = b.z;
        }
        const bool never = sm[L.octime + e * 3 + 0] < A.os_last_contact_thr && sm[L.octime + e * 3 + 1] < A.os_current_contact_thr;
        const float term_scale = A.obs_terms[tobj].scale;
#pragma unroll
        for (int k = 0; k < 13; ++k) {
          const float val = __fmul_rn(never ? cst[k] : st[k], A.os_scale[k]);
          nv[(grp * kEnvs + e) * dps + jb + k] = __fmul_rn(val, term_scale);
        }
      }
    }
    // publish any(non_zero_cmd) for the reward pass of the NEXT step (it sees the same command tensor)
    if (A.any_flag_ws && tid < nvalid) {
      const float* c = sm + L.cmd + tid * 3;
      if (sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) > 0.f) A.any_flag_ws[(step + 1) & 1] = step + 1;
    }
  }
  __syncthreads();

  // ----------------------------------------------------------------------------------------- stage 2: terminations
  if (do_rew && tid < nvalid) {
    const int e = tid, n = e0 + e;
    bool terminated = false, timed_out = false;
    for (int t = 0; t < A.num_termination_terms; ++t) {
      const LtTerminationTerm& tt = A.termination_terms[t];
      bool m = false;
      switch (tt.kind) {
        case LT_TK_TIME_OUT: m = A.episode_length_buf[n] >= A.max_episode_length; break;
        case LT_TK_BAD_ORIENTATION: m = fabsf(acosf(-sm[L.grav + e * 3 + 2])) > tt.p[0]; break;
        case LT_TK_ROOT_HEIGHT: m = sm[L.pos + e * 3 + 2] < tt.p[0]; break;
        case LT_TK_ILLEGAL_CONTACT:
          for (int k = 0; k < tt.num_ids; ++k) m = m || sm[L.fmax + e * S + tt.body_ids[k]] > tt.p[0];
          break;
        case LT_TK_OBJECT_BELOW_ROBOT: m = sm[L.opos + e * 3 + 2] < sm[L.pos + e * 3 + 2]; break;
        case LT_TK_BAD_ROLL: m = fabsf(asinf(sm[L.ograv + e * 3 + 1])) > tt.p[0]; break;
      }
      if (A.term_masks) A.term_masks[(size_t)t * A.N + n] = m;
      if (tt.time_out) timed_out = timed_out || m; else terminated = terminated || m;
    }
    s_term[e] = terminated;
    s_done[e] = terminated || timed_out;
    A.terminated[n] = terminated;
    A.time_outs[n] = timed_out;
    A.dones[n] = terminated || timed_out;
    if (A.auto_reset && (terminated || timed_out)) {  // history of a reset env is refilled by its next observation
      if (do_obs) s_fill[e] = 1;
      else if (A.obs_fill) A.obs_fill[n] = 1;
    }
  }
  __syncthreads();

  // --------------------------------------------------------------------------------------- stage 3a: rewards (warp 0)
  if (warp == 0) {
    if (!do_rew || lane >= nvalid) return;
    const int e = lane, n = e0 + e;
    const float dt = A.step_dt;
    const Vec3 cmd = ld3(sm + L.cmd + e * 3), vb = ld3(sm + L.linb + e * 3), wb = ld3(sm + L.angb + e * 3);
    const Vec3 grav = ld3(sm + L.grav + e * 3), pos = ld3(sm + L.pos + e * 3);
    const float cmd_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y + cmd.z * cmd.z);
    const bool nz = cmd_norm > 0.f;
    const float* fmx = sm + L.fmax + e * S;
    const float* fpos = sm + L.fpos + e * 12;
    const float* fvel = sm + L.fvel + e * 12;
    float foot_speed[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) foot_speed[k] = sqrtf(fvel[3 * k] * fvel[3 * k] + fvel[3 * k + 1] * fvel[3 * k + 1]);
    // object-relative quantities shared by several terms
    Quat q = {1.f, 0.f, 0.f, 0.f};
    Vec3 rel_pos_w = {0, 0, 0}, rel_pos = {0, 0, 0}, rel_vel = {0, 0, 0}, rel_ang = {0, 0, 0}, g_obj = {0, 0, 0};
    if (has_obj) {
      q = ld4(sm + L.quat + e * 4);
      rel_pos_w = sub3(ld3(sm + L.opos + e * 3), pos);
      rel_pos = rot_inv(q, rel_pos_w);
      rel_vel = rot_inv(q, sub3(ld3(sm + L.olin + e * 3), ld3(sm + L.linw + e * 3)));
      rel_ang = rot_inv(q, sub3(ld3(sm + L.oang + e * 3), ld3(sm + L.angw + e * 3)));
      g_obj = rot_inv(q, rot(ld4(sm + L.oquat + e * 4), ld3(sm + L.ograv + e * 3)));
    }

    float reward = 0.f;
    for (int i = 0; i < A.num_reward_terms; ++i) {
      const LtRewardTerm& rt = A.reward_terms[i];
      if (rt.weight == 0.f) {  // [IL] RewardManager.compute skips zero-weight terms
        if (A.step_reward) A.step_reward[(size_t)n * A.num_reward_terms + i] = 0.f;
        continue;
      }
      float raw = 0.f;
      switch (rt.kind) {
        case LT_RK_ALIVE: raw = s_term[e] ? 0.f : 1.f; break;
        case LT_RK_TRACK_LIN_VEL_XY: {
          const float dx = cmd.x - vb.x, dy = cmd.y - vb.y;
          raw = expf(-sqrtf(dx * dx + dy * dy) / rt.p[0]);
        } break;
        case LT_RK_TRACK_ANG_VEL_Z: raw = expf(-fabsf(cmd.z - wb.z) / rt.p[0]); break;
        case LT_RK_FOOT_SLIP: {
#pragma unroll
          for (int k = 0; k < 4; ++k) raw += (fmx[A.feet_sensor_ids[k]] > rt.p[0] ? 1.f : 0.f) * foot_speed[k];
        } break;
        case LT_RK_FOOT_DRAG: {
#pragma unroll
          for (int k = 0; k < 4; ++k) raw += (fpos[3 * k + 2] <= rt.p[0] && foot_speed[k] > rt.p[1]) ? 1.f : 0.f;
        } break;
        case LT_RK_GAIT: {
          const LtGaitParams& gp = A.gait;
          const float th = gp.judge_time_threshold;
          GaitRegs g;
          const float* gs = sm + L.gait + e * 24;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            g.lsa[k] = gs[k]; g.lsc[k] = gs[4 + k]; g.vla[k] = gs[8 + k];
            g.sz[k] = gs[16 + k] != 0.f; g.vpc[k] = gs[20 + k] != 0.f;
          }
          g.last_cmd[0] = gs[12]; g.last_cmd[1] = gs[13]; g.last_cmd[2] = gs[14]; g.steps = gs[15];
          float a[4], c[4], la[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) { a[k] = sm[L.air + e * 4 + k]; c[k] = sm[L.con + e * 4 + k]; la[k] = sm[L.lair + e * 4 + k]; }
          bool any_nz = true;
          if (A.any_nonzero_cmd_override >= 0) any_nz = A.any_nonzero_cmd_override != 0;
          else if (A.any_flag_ws) any_nz = A.any_flag_ws[step & 1] == step;
          gait_update(g, a, c, la, cmd, nz, any_nz, th);
          // task performance score (rewards.py:202-216 / 372-392)
          float score = 0.f;
          if (gp.encourage_symmetricity) {
            const float dx = cmd.x - vb.x, dy = cmd.y - vb.y;
            const float e_lin = nz ? sqrtf(dx * dx + dy * dy) : 0.f;
            const float e_ang = nz ? fabsf(cmd.z - wb.z) : 0.f;
            score = (expf(-e_lin / gp.vel_tracking_exp_sigma) + expf(-e_ang / gp.vel_tracking_exp_sigma)) / 2.f;
            if (gp.with_object) {
              const Vec3 r = rot_inv(yaw_quat(yaw_of(q)), rel_pos_w);
              const float bx = clampf(1.f - fabsf(r.x) / gp.obj_x_max, 0.f, 1.f);
              const float by = clampf(1.f - fabsf(r.y) / gp.obj_y_max, 0.f, 1.f);
              score = clampf((score * 2.f + (bx + by) / 2.f) / 3.f, 0.f, 1.f);
            }
          }
          const float sync = (gait_sync(gp, g, a, c, 0, score) + gait_sync(gp, g, a, c, 1, score)) / 2.f;
          const float asyn = (gait_async(gp, a, c, 0, 2) + gait_async(gp, a, c, 1, 3) + gait_async(gp, a, c, 0, 3) + gait_async(gp, a, c, 2, 1)) / 4.f;
          const float stepping = (sync + asyn) / 2.f;
          const float stance = ((c[0] > th && c[1] > th && c[2] > th && c[3] > th) ? 1.f : 0.f) * gp.stance_rwd_scale;
          raw = nz ? stepping : stance;
          // write the state back (zeroed when this env is being reset: the manager's reset(env_ids) follows, rewards.py:107-114)
          const bool rst = A.auto_reset && s_done[e];
          const LtGaitState& G = A.gait_state;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            G.last_step_current_air_time[n * 4 + k] = rst ? 0.f : g.lsa[k];
            G.last_step_current_contact_time[n * 4 + k] = rst ? 0.f : g.lsc[k];
            G.valid_last_air_time[n * 4 + k] = rst ? 0.f : g.vla[k];
            G.swinging_in_zero_cmd[n * 4 + k] = rst ? 0 : (g.sz[k] ? 1 : 0);
            G.valid_previous_contact[n * 4 + k] = rst ? 0 : (g.vpc[k] ? 1 : 0);
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) G.last_velocity_cmd[n * 3 + k] = rst ? 0.f : g.last_cmd[k];
          G.step_from_changing_cmd[n] = rst ? 0.f : g.steps;
        } break;
        case LT_RK_BASE_HEIGHT: { const float d = pos.z - rt.p[0]; raw = d * d; } break;
        case LT_RK_BASE_Z_VEL: raw = vb.z * vb.z; break;
        case LT_RK_BASE_RP_ANGLE: raw = grav.x * grav.x + grav.y * grav.y; break;
        case LT_RK_BASE_RP_VEL: raw = fabsf(wb.x) + fabsf(wb.y); break;
        case LT_RK_JOINT_POS_LIMIT: {
          const float* qq = sm + L.q + e * J; const float* lim = sm + L.lim + e * 2 * J;
          for (int j = 0; j < J; ++j) raw += -fminf(qq[j] - lim[2 * j], 0.f) + fmaxf(qq[j] - lim[2 * j + 1], 0.f);
        } break;
        case LT_RK_JOINT_POS: {
          const float* qq = sm + L.q + e * J; const float* q0 = sm + L.q0 + e * J;
          float s = 0.f;
          for (int j = 0; j < J; ++j) { const float d = qq[j] - q0[j]; s += d * d; }
          const float dev = sqrtf(s);
          const float bv = sqrtf(vb.x * vb.x + vb.y * vb.y);
          raw = (cmd_norm > 0.f || bv > rt.p[1]) ? dev : rt.p[0] * dev;
        } break;
        case LT_RK_JOINT_ACC: case LT_RK_JOINT_VEL: case LT_RK_JOINT_TORQUE: {
          const float* x = sm + (rt.kind == LT_RK_JOINT_ACC ? L.qdd : (rt.kind == LT_RK_JOINT_VEL ? L.qd : L.tau)) + e * J;
          float s = 0.f;
          for (int j = 0; j < J; ++j) s += x[j] * x[j];
          raw = sqrtf(s);
        } break;
        case LT_RK_ACTION_RATE: {
          const float* a = sm + L.act + e * J; const float* pa = sm + L.pact + e * J;
          for (int j = 0; j < J; ++j) { const float d = a[j] - pa[j]; raw += d * d; }
        } break;
        case LT_RK_THIGH_CALF_COLLISION:
          for (int k = 0; k < A.num_thigh_calf; ++k) raw += fmx[A.thigh_calf_sensor_ids[k]] > rt.p[0] ? 1.f : 0.f;
          break;
        case LT_RK_OBJ_XY_POS:
          raw = sqrtf(rel_pos_w.x * rel_pos_w.x + rel_pos_w.y * rel_pos_w.y);
          if (rt.p[0] != 0.f) raw *= (cmd_norm > 0.f) ? 1.f : 0.f;
          break;
        case LT_RK_OBJ_XY_VEL: raw = rel_vel.x * rel_vel.x + rel_vel.y * rel_vel.y; break;
        case LT_RK_OBJ_LOSE_CONTACT: raw = (sm[L.octime + e * 3 + 0] > 0.f && sm[L.octime + e * 3 + 2] > 0.f) ? 1.f : 0.f; break;
        case LT_RK_OBJ_Z_VEL: raw = rel_vel.z * rel_vel.z; break;
        case LT_RK_OBJ_RP_ANGLE: raw = g_obj.x * g_obj.x + g_obj.y * g_obj.y; break;
        case LT_RK_OBJ_RP_VEL: raw = fabsf(rel_ang.x) + fabsf(rel_ang.y); break;
        case LT_RK_OBJ_ROLL_ANGLE: raw = g_obj.y * g_obj.y; break;
        case LT_RK_OBJ_ROLL_VEL: raw = rel_ang.x * rel_ang.x; break;
        case LT_RK_OBJ_YAW: {  // rewards.py:545-567
          const Quat qr = yaw_quat(yaw_of(q)), qo = yaw_quat(yaw_of(ld4(sm + L.oquat + e * 4)));
          float d = yaw_of(quat_mul(quat_inv(qr), qo));
          const float pi = 3.14159274101257324f;  // float32(torch.pi)
          if (d > pi) d -= 2.f * pi;
          if (d > 0.5f * pi) d -= pi;
          if (d <= -0.5f * pi) d += pi;
          raw = d * d;
          if (rt.p[0] != 0.f) raw *= (cmd_norm > 0.f) ? 1.f : 0.f;
        } break;
        case LT_RK_OBJ_DANGER: {  // rewards.py:569-594
          bool bad = fabsf(rel_pos.x) > rt.p[0];
          bad = bad || fabsf(rel_pos.y) > rt.p[1];
          bad = bad || rel_pos.z < rt.p[2];
          if (rt.p[3] >= 0.f) bad = bad || fabsf(acosf(-sm[L.ograv + e * 3 + 2])) > rt.p[3] * 3.14159265358979323846f / 180.f;
          if (rt.p[4] >= 0.f) bad = bad || sqrtf(rel_vel.x * rel_vel.x + rel_vel.y * rel_vel.y) > rt.p[4];
          raw = bad ? 1.f : 0.f;
        } break;
        default: break;
      }
      // [IL] RewardManager.compute: value = f * weight * dt ; reward += value ; episode_sums += value ; step_reward = value / dt
      const float value = __fmul_rn(__fmul_rn(raw, rt.weight), dt);
      reward = __fadd_rn(reward, value);
      if (A.term_raw) A.term_raw[(size_t)i * A.N + n] = raw;
      if (A.step_reward) A.step_reward[(size_t)n * A.num_reward_terms + i] = __fdiv_rn(value, dt);
      if (A.episode_sums) {
        float* es = A.episode_sums + (size_t)i * A.N + n;
        const float total = __fadd_rn(*es, value);
        if (A.auto_reset && s_done[e]) {
          if (A.episode_log_sums) atomicAdd(A.episode_log_sums + i, total);
          *es = 0.f;
        } else {
          *es = total;
        }
      }
    }
    A.reward[n] = reward;
    if (A.auto_reset && s_done[e] && A.episode_log_sums) atomicAdd(A.episode_log_sums + A.num_reward_terms, 1.0f);
    return;
  }

  // ------------------------------------------------------------------------- stage 3b: observation history (warps 1..3)
  if (!do_obs) return;
  {
    const int D = L.D, dps = L.dps;
    const int otid = tid - 32;
    const int per_group = nvalid * D;
    const int total = 2 * per_group;
    const float* nv = sm + L.newobs;
    float vals[kMaxObsRegs];
#pragma unroll
    for (int r = 0; r < kMaxObsRegs; ++r) {
      const int idx = otid + r * kObsThreads;
      if (idx < total) {
        const int grp = idx >= per_group, rem = idx - grp * per_group;
        const int e = rem / D, k = rem - e * D;
        const int entry = s_map[k];
        const int j = entry & 0xff, d = (entry >> 8) & 0xff;
        const bool newest = (entry >> 16) & 1;
        const float* in = grp ? A.critic_obs_in : A.policy_obs_in;
        if (newest || s_fill[e] || !in) vals[r] = nv[(grp * kEnvs + e) * dps + j];
        else vals[r] = __ldcs(in + (size_t)(e0 + e) * D + k + d);
      }
    }
    named_barrier_sync(1, kObsThreads);  // in/out may alias: every load of the block precedes its stores
#pragma unroll
    for (int r = 0; r < kMaxObsRegs; ++r) {
      const int idx = otid + r * kObsThreads;
      if (idx < total) {
        const int grp = idx >= per_group, rem = idx - grp * per_group;
        float* out = grp ? A.critic_obs_out : A.policy_obs_out;
        if (out) __stcs(out + (size_t)e0 * D + rem, vals[r]);
      }
    }
    if (A.obs_fill && otid < nvalid) A.obs_fill[e0 + otid] = 0;
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
    for (int i = 0; i < a->num_reward_terms; ++i) {
      const int k = a->reward_terms[i].kind;
      if (k < 0 || k >= LT_RK_COUNT) return LT_ERR_INVALID_ARG;
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
  int off = 0;
  auto take = [&](int per_env) { const int o = off; off += kEnvs * per_env; return o; };
  L.cmd = take(3); L.pos = take(3); L.linb = take(3); L.angb = take(3); L.grav = take(3);
  L.q = take(J); L.qd = take(J); L.qdd = take(J); L.tau = take(J); L.q0 = take(J); L.qd0 = take(J); L.lim = take(2 * J);
  L.act = take(J); L.pact = take(J);
  L.force = take(do_rew ? H * S * 3 : 0);
  L.air = take(4); L.con = take(4); L.lair = take(4); L.fpos = take(12); L.fvel = take(12);
  L.quat = take(4); L.linw = take(3); L.angw = take(3); L.opos = take(3); L.oquat = take(4); L.olin = take(3); L.oang = take(3);
  L.ograv = take(3); L.octime = take(3);
  L.fmax = take(S); L.gait = take(24);
  L.newobs = take(2 * dps);
  L.map = off; off += L.D;
  L.total = off;
  const size_t smem = (size_t)L.total * sizeof(float);
  if (smem > 48 * 1024) return LT_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (do_rew && a->any_nonzero_cmd_override == -2) {
    if (!a->any_flag_ws) return LT_ERR_INVALID_ARG;
    any_nonzero_cmd_kernel<<<1, 1024, 0, st>>>(a->command, a->N, a->any_flag_ws, (int)a->offset);
    int rc = lt::check_launch();
    if (rc != LT_OK) return rc;
  }
  const int grid = (int)lt::ceil_div(a->N, kEnvs);
  LtMdpArgs args = *a;
  if (args.any_nonzero_cmd_override == -2) args.any_nonzero_cmd_override = -1;
  mdp_step_kernel<<<grid, kThreads, smem, st>>>(args, L);
  return lt::check_launch();
}

extern "C" int lt_mdp_reset(const LtGaitState* g, float* episode_sums, int num_reward_terms, const uint8_t* mask, int N, void* stream) {
  if (!g || !mask || N <= 0) return LT_ERR_INVALID_ARG;
  mdp_reset_kernel<<<(unsigned)lt::ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(*g, episode_sums, num_reward_terms, mask, N);
  return lt::check_launch();
}
