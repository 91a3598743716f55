// K2: binary taxel synthesis from back-mounted contact forces (+ the tactile delay line).
// Replaces reference locotouch/mdp/observations.py:154-159 (get_original_signals), :166-199 (get_normal_forces) and
// :281-308 (BinaryTactileSignals.__call__) -- 385 ATen launches per step in the reference -- and
// locotouch/distill/tactile_recorder.py:4-34 (TactileRecorder).
//
// One warp per env.  The 221 taxels of an env are walked in 7 rounds of 32 lanes: each lane loads its taxel's link
// quaternion with one 128-bit access, the world force, the per-(env,taxel) threshold, rotates the force into the
// taxel frame, thresholds it (strict '>'), applies dropout / addition, and the warp ballots the 32 verdicts into one
// word of the packed bitmap.  The fp32 [N, 2*221] tensor the student consumes is written from the same registers;
// the delay line is kept on the packed words (28 B/env/frame instead of 1768 B).
// The normal force is evaluated in the reference's operation order with contraction disabled so that the comparison
// against the threshold sees the same fp32 value as the torch expression.
// Algorithmic traffic per env-step: 221 x (16 + 12 + 4) B read + 442 x 4 B written = 8840 B.
#include "lt_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxWords = 32;  // T <= 1024

// z component of quat_apply_inverse(q, v) ([IL] isaaclab.utils.math): v - w*t + xyz x t, t = 2 (xyz x v)
__device__ __forceinline__ float rotate_inverse_z(float4 q /*w,x,y,z*/, float vx, float vy, float vz) {
  const float w = q.x, x = q.y, y = q.z, z = q.w;
  const float tx = __fmul_rn(__fsub_rn(__fmul_rn(y, vz), __fmul_rn(z, vy)), 2.0f);
  const float ty = __fmul_rn(__fsub_rn(__fmul_rn(z, vx), __fmul_rn(x, vz)), 2.0f);
  const float tz = __fmul_rn(__fsub_rn(__fmul_rn(x, vy), __fmul_rn(y, vx)), 2.0f);
  const float cz = __fsub_rn(__fmul_rn(x, ty), __fmul_rn(y, tx));
  return __fadd_rn(__fsub_rn(vz, __fmul_rn(w, tz)), cz);
}

constexpr int kMaxDelay = 8;

// R = number of 32-taxel rounds held in registers (R == 0: generic loop for any T).  With R fixed every global load of
// the env (quaternion, force, threshold, uniforms, old delay-ring words) is issued before the first ballot, so a warp pays
// one DRAM round trip instead of one per round.
template <int R>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) taxel_kernel(const LtTaxelArgs a) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= a.N) return;  // warp-uniform
  const int T = a.T, words = (T + 31) >> 5;
  const float4* quat = reinterpret_cast<const float4*>(a.body_quat_w) + (size_t)n * a.quat_num_bodies + a.quat_body_offset;
  const float* force = a.net_forces_w + (size_t)n * T * 3;
  const size_t row = (size_t)n * T;
  const bool explicit_u = a.u_drop || a.u_add;
  // ---- delay line state, fetched up front (tactile_recorder.py:25-34 on the packed words)
  uint32_t old_ring[kMaxDelay - 1];
  bool first = false;
  int slot = 0;
  if (a.delay_ring) {
    const uint32_t* ring = a.delay_ring + (size_t)n * a.max_delay * words;
#pragma unroll
    for (int k = 0; k < kMaxDelay - 1; ++k) old_ring[k] = (k < a.max_delay - 1 && lane < words) ? ring[k * words + lane] : 0u;
    first = a.delay_first[n] != 0 || (a.delay_reset != nullptr && a.delay_reset[n] != 0);
    slot = (int)a.delay_steps[n];
  }
  const uint64_t rng_offset = a.offset + (a.offset_base ? (uint64_t)*a.offset_base : 0ull);
  uint32_t my_word = 0;  // lane r keeps word r

  auto decide = [&](int t, float4 q, float fx, float fy, float fz, float thr, float ud, float ua) -> bool {
    const float fn = -rotate_inverse_z(q, fx, fy, fz);  // observations.py:156-158
    const bool original = fn > thr;                     // observations.py:159 (strict)
    bool contact = original;
    if (a.p_drop > 0.f) contact = contact && !(ud < a.p_drop);  // observations.py:172-176
    if (a.p_add > 0.f) contact = contact || (ua < a.p_add);     // observations.py:180-185
    if (a.normal_forces) a.normal_forces[row + t] = fn;
    if (a.original_contact) a.original_contact[row + t] = original ? 1 : 0;
    if (a.signal) {
      const float s = contact ? 1.0f : 0.0f;
      float* out = a.signal + (size_t)n * 2 * T;
      __stcs(out + t, s);      // channel 0
      __stcs(out + T + t, s);  // channel 1 (observations.py:308: two identical channels)
    }
    return contact;
  };

  if constexpr (R > 0) {
    float4 q[R];
    float fx[R], fy[R], fz[R], thr[R], ud[R], ua[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = 32 * r + lane;
      q[r] = make_float4(1.f, 0.f, 0.f, 0.f);
      fx[r] = fy[r] = fz[r] = 0.f;
      thr[r] = 0.f;
      ud[r] = ua[r] = 1.f;
      if (t < T) {
        q[r] = __ldcs(quat + t);
        fx[r] = __ldcs(force + 3 * t); fy[r] = __ldcs(force + 3 * t + 1); fz[r] = __ldcs(force + 3 * t + 2);
        thr[r] = __ldcs(a.thresholds + row + t);
        if (explicit_u) {
          if (a.u_drop) ud[r] = __ldcs(a.u_drop + row + t);
          if (a.u_add) ua[r] = __ldcs(a.u_add + row + t);
        }
      }
    }
    if (!explicit_u) {
#pragma unroll
      for (int r = 0; r < R; r += 2) {
        const uint4 rnd = lt::Philox::gen(a.seed, rng_offset, (uint32_t)n, (uint32_t)(32 * (r >> 1) + lane));
        ud[r] = lt::Philox::u01(rnd.x); ua[r] = lt::Philox::u01(rnd.y);
        if (r + 1 < R) { ud[r + 1] = lt::Philox::u01(rnd.z); ua[r + 1] = lt::Philox::u01(rnd.w); }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = 32 * r + lane;
      const bool contact = (t < T) && decide(t, q[r], fx[r], fy[r], fz[r], thr[r], ud[r], ua[r]);
      const uint32_t word = __ballot_sync(LT_FULL_MASK, contact);
      if (lane == r) my_word = word;
    }
  } else {
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int r = 0; r < words; ++r) {
      const int t = 32 * r + lane;
      bool contact = false;
      if (t < T) {
        const float4 q = __ldcs(quat + t);
        const float fx = __ldcs(force + 3 * t), fy = __ldcs(force + 3 * t + 1), fz = __ldcs(force + 3 * t + 2);
        const float thr = __ldcs(a.thresholds + row + t);
        float ud, ua;
        if (explicit_u) {
          ud = a.u_drop ? __ldcs(a.u_drop + row + t) : 1.0f;
          ua = a.u_add ? __ldcs(a.u_add + row + t) : 1.0f;
        } else {
          if ((r & 1) == 0) rnd = lt::Philox::gen(a.seed, rng_offset, (uint32_t)n, (uint32_t)(32 * (r >> 1) + lane));
          ud = lt::Philox::u01((r & 1) ? rnd.z : rnd.x);
          ua = lt::Philox::u01((r & 1) ? rnd.w : rnd.y);
        }
        contact = decide(t, q, fx, fy, fz, thr, ud, ua);
      }
      const uint32_t word = __ballot_sync(LT_FULL_MASK, contact);
      if (lane == r) my_word = word;
    }
  }
  if (a.packed && lane < words) a.packed[(size_t)n * words + lane] = my_word;

  if (a.delay_ring) {
    // shifted ring: ring'[0] = new, ring'[k] = first ? new : ring[k-1]; output slot = ring'[delay]  (all from registers)
    uint32_t* ring = a.delay_ring + (size_t)n * a.max_delay * words;
    uint32_t delayed = my_word;
    if (lane < words) {
      ring[lane] = my_word;
#pragma unroll
      for (int k = 1; k < kMaxDelay; ++k) {
        if (k < a.max_delay) {
          const uint32_t v = first ? my_word : old_ring[k - 1];
          ring[k * words + lane] = v;
          if (k == slot) delayed = v;
        }
      }
    }
    if (lane == 0) a.delay_first[n] = 0;
    if (a.delayed_signal) {
      float* out = a.delayed_signal + (size_t)n * 2 * T;
      for (int r = 0; r < words; ++r) {
        const uint32_t w = __shfl_sync(LT_FULL_MASK, delayed, r);
        const int t = 32 * r + lane;
        if (t < T) {
          const float s = (w >> lane) & 1u ? 1.0f : 0.0f;
          __stcs(out + t, s);
          __stcs(out + T + t, s);
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Force-valued encodings (reference observations.py:166-237, classes :311-429): one warp per env, up to kMaxRounds x 32 taxels
// held in registers so that the per-env min / max of the normalised forces (a warp reduction) and the second pass over the
// taxels need no re-read.  Expression order follows the torch code (file built with -fmad=false).
constexpr int kMaxRounds = 8;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) taxel_force_kernel(const LtTaxelForceArgs a) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= a.N) return;  // warp-uniform
  const int T = a.T, rounds = (T + 31) >> 5;
  const float4* quat = reinterpret_cast<const float4*>(a.body_quat_w) + (size_t)n * a.quat_num_bodies + a.quat_body_offset;
  const float* force = a.net_forces_w + (size_t)n * T * 3;
  const size_t row = (size_t)n * T;
  const uint64_t rng_offset = a.offset + (a.offset_base ? (uint64_t)*a.offset_base : 0ull);
  float norm[kMaxRounds], fn_keep[kMaxRounds];
  bool con[kMaxRounds];
  float vmin = 3.0e38f, vmax = -3.0e38f;
#pragma unroll
  for (int r = 0; r < kMaxRounds; ++r) {
    norm[r] = 0.f; fn_keep[r] = 0.f; con[r] = false;
    const int t = 32 * r + lane;
    if (r < rounds && t < T) {
      const float4 q = __ldcs(quat + t);
      const float thr = __ldcs(a.thresholds + row + t);
      float fn = -rotate_inverse_z(q, __ldcs(force + 3 * t), __ldcs(force + 3 * t + 1), __ldcs(force + 3 * t + 2));  // :156-158
      // the seven uniforms of this taxel: explicit tensors (parity) or two Philox draws keyed by (env, taxel)
      float u[7];
      if (a.u[0]) {
#pragma unroll
        for (int k = 0; k < 7; ++k) u[k] = a.u[k] ? __ldcs(a.u[k] + row + t) : 1.0f;
      } else {
        const uint4 r0 = lt::Philox::gen(a.seed, rng_offset, (uint32_t)n, (uint32_t)(2 * t));
        const uint4 r1 = lt::Philox::gen(a.seed, rng_offset, (uint32_t)n, (uint32_t)(2 * t + 1));
        u[0] = lt::Philox::u01(r0.x); u[1] = lt::Philox::u01(r0.y); u[2] = lt::Philox::u01(r0.z); u[3] = lt::Philox::u01(r0.w);
        u[4] = lt::Philox::u01(r1.x); u[5] = lt::Philox::u01(r1.y); u[6] = lt::Philox::u01(r1.z);
      }
      bool contact = fn > thr;  // :159 (strict)
      if (a.p_drop > 0.f && contact && u[0] < a.p_drop) {  // :170-176 dropped contact: force in [0, threshold)
        fn = u[1] * thr;
        contact = false;
      }
      if (a.p_add > 0.f && !contact && u[2] < a.p_add) {  // :180-185 added contact: force in [threshold, 1.2 threshold)
        fn = thr * (1.0f + 0.2f * u[3]);
        contact = true;
      }
      if (a.add_force_noise) {  // :188-193
        if (contact) fn = fn * (1.0f + (u[4] * a.force_noise_range + a.force_noise_min));
        fn = fmaxf(fn, 0.0f);
        if (contact && fn < thr) fn = thr * (1.0f + 0.2f * u[5]);
      }
      const float nf = fminf(fmaxf(fn / a.maximal_force, 0.0f), 1.0f);  // :203
      const float valid = contact ? nf : 0.0f;                          // :208
      vmin = fminf(vmin, valid);
      vmax = fmaxf(vmax, valid);
      norm[r] = nf; fn_keep[r] = u[6]; con[r] = contact;  // fn_keep carries the level-noise uniform to the second pass
      if (a.normal_forces) a.normal_forces[(size_t)n * a.out_stride + t] = fn;
      if (a.contact) a.contact[(size_t)n * a.out_stride + t] = contact ? 1.0f : 0.0f;
      if (a.normalized) a.normalized[(size_t)n * a.out_stride + t] = nf;
    }
  }
  if (!a.minmax && !a.discretized) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(LT_FULL_MASK, vmin, o));
    vmax = fmaxf(vmax, __shfl_xor_sync(LT_FULL_MASK, vmax, o));
  }
  const float range = (vmax - vmin) > 0.0f ? vmax - vmin : 1.0f;  // :215-217
#pragma unroll
  for (int r = 0; r < kMaxRounds; ++r) {
    const int t = 32 * r + lane;
    if (r < rounds && t < T) {
      const float valid = con[r] ? norm[r] : 0.0f;
      const float mm = fminf(fmaxf((valid - vmin) / range, 0.0f), 1.0f);  // :220-221
      if (a.minmax) a.minmax[(size_t)n * a.out_stride + t] = mm;
      if (a.discretized) {
        float d = rintf(mm / a.level_bin);  // torch.round: half to even (:229)
        if (a.add_level_noise) d = d + (fn_keep[r] * a.level_noise_range + a.level_noise_min);
        d = d * a.level_bin;
        d = fminf(fmaxf(d, 0.0f), 1.0f);
        a.discretized[(size_t)n * a.out_stride + t] = con[r] ? d : 0.0f;  // :235
      }
    }
  }
}

// Generic fp32 delay line, one warp per env (the `first` flag is read and cleared by the same warp).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
tactile_delay_kernel(float* __restrict__ ring, uint8_t* __restrict__ first, const int64_t* __restrict__ delay_steps,
                     const float* __restrict__ signal, float* __restrict__ out, int N, int max_delay, int D) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= N) return;
  const bool fill = first[n] != 0;
  const int slot = (int)delay_steps[n];
  float* r = ring + (size_t)n * max_delay * D;
  for (int d = lane; d < D; d += 32) {
    const float x = __ldcs(signal + (size_t)n * D + d);
    for (int k = max_delay - 1; k >= 1; --k) r[(size_t)k * D + d] = fill ? x : r[(size_t)(k - 1) * D + d];
    r[d] = x;
    out[(size_t)n * D + d] = r[(size_t)slot * D + d];
  }
  __syncwarp();
  if (lane == 0) first[n] = 0;
}

}  // namespace

extern "C" int lt_taxel_synth(const LtTaxelArgs* a, void* stream) {
  if (!a || a->N <= 0 || a->T <= 0 || a->T > 32 * kMaxWords) return LT_ERR_INVALID_ARG;
  if (!a->body_quat_w || !a->net_forces_w || !a->thresholds) return LT_ERR_INVALID_ARG;
  if (((uintptr_t)a->body_quat_w & 15) != 0) return LT_ERR_INVALID_ARG;
  if (a->quat_body_offset < 0 || a->quat_body_offset + a->T > a->quat_num_bodies) return LT_ERR_INVALID_ARG;
  if (a->delay_ring && (!a->delay_first || !a->delay_steps || a->max_delay <= 0 || a->max_delay > kMaxDelay)) return LT_ERR_INVALID_ARG;
  const int grid = (int)lt::ceil_div(a->N, kWarpsPerBlock);
  if (a->T > 192 && a->T <= 224)
    taxel_kernel<7><<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(*a);  // 17 x 13 = 221 taxels
  else
    taxel_kernel<0><<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(*a);
  return lt::check_launch();
}

extern "C" int lt_taxel_forces(const LtTaxelForceArgs* a, void* stream) {
  if (!a || a->N <= 0 || a->T <= 0 || a->T > 32 * kMaxRounds) return LT_ERR_INVALID_ARG;
  if (!a->body_quat_w || !a->net_forces_w || !a->thresholds || a->out_stride < a->T) return LT_ERR_INVALID_ARG;
  if (((uintptr_t)a->body_quat_w & 15) != 0) return LT_ERR_INVALID_ARG;
  if (a->quat_body_offset < 0 || a->quat_body_offset + a->T > a->quat_num_bodies) return LT_ERR_INVALID_ARG;
  if (!(a->maximal_force > 0.f) || !(a->level_bin > 0.f)) return LT_ERR_INVALID_ARG;
  if (!a->contact && !a->normal_forces && !a->normalized && !a->minmax && !a->discretized) return LT_ERR_INVALID_ARG;
  const int grid = (int)lt::ceil_div(a->N, kWarpsPerBlock);
  taxel_force_kernel<<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(*a);
  return lt::check_launch();
}

extern "C" int lt_tactile_delay(float* ring, uint8_t* first, const int64_t* delay_steps, const float* signal, float* out, int N,
                                int max_delay, int D, void* stream) {
  if (!ring || !first || !delay_steps || !signal || !out || N <= 0 || max_delay <= 0 || D <= 0) return LT_ERR_INVALID_ARG;
  const int grid = (int)lt::ceil_div(N, kWarpsPerBlock);
  tactile_delay_kernel<<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(ring, first, delay_steps, signal, out, N, max_delay, D);
  return lt::check_launch();
}
