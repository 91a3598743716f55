// K3 / K5: the per-step bookkeeping of the PPO rollout.
//  * lt_act_sample  -- Normal(mu, sigma).sample() + log_prob().sum(-1) of reference
//                      loco_rl/loco_rl/modules/actor_critic.py:105-123 as driven by PPO.act (algorithms/ppo.py:129-141);
//                      outputs are written straight into the RolloutStorage slot (no staging tensors, no copy_ kernels).
//  * lt_store_step  -- time-out bootstrap (ppo.py:162-165) fused with RolloutStorage.add_transitions
//                      (storage/rollout_storage.py:80-107).
//  * lt_gather_rows -- the nine advanced-index gathers of RolloutStorage.mini_batch_generator
//                      (rollout_storage.py:221-231) as one launch.
#include "lt_common.cuh"

namespace {

constexpr float kHalfLog2Pi = 0.91893853320467274178f;

// 4 lanes per env row (A multiple of 4): float4 chunk per lane, lane-group reduction for the log-prob.
__global__ void __launch_bounds__(256)
act_sample_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, const float* __restrict__ eps,
                  float* __restrict__ actions, float* __restrict__ logp, float* __restrict__ mu_out,
                  float* __restrict__ sigma_out, int N, int A, uint64_t seed, uint64_t offset, const int64_t* __restrict__ offset_base) {
  if (offset_base) offset += (uint64_t)*offset_base;
  const int chunks = A >> 2;
  const int sub = threadIdx.x & 3;
  const int base = (blockIdx.x * blockDim.x) >> 2;
  const int n = base + (threadIdx.x >> 2);
  const bool valid = n < N;
  float lp = 0.f;
  for (int ch = sub; ch < chunks; ch += 4) {
    if (!valid) continue;
    const size_t off = (size_t)n * A + 4 * ch;
    const float4 m = __ldg(reinterpret_cast<const float4*>(mu + off));
    const float4 s = __ldg(reinterpret_cast<const float4*>(sigma + 4 * ch));
    float4 e;
    if (eps) {
      e = __ldcs(reinterpret_cast<const float4*>(eps + off));
    } else {
      const uint4 r = lt::Philox::gen(seed, offset, (uint32_t)n, (uint32_t)ch);
      const float2 a = lt::Philox::normal2(r.x, r.y), b = lt::Philox::normal2(r.z, r.w);
      e = make_float4(a.x, a.y, b.x, b.y);
    }
    // torch.normal(mean, std): mean + std * eps
    const float4 a = make_float4(__fadd_rn(m.x, __fmul_rn(s.x, e.x)), __fadd_rn(m.y, __fmul_rn(s.y, e.y)),
                                 __fadd_rn(m.z, __fmul_rn(s.z, e.z)), __fadd_rn(m.w, __fmul_rn(s.w, e.w)));
    const float mm[4] = {m.x, m.y, m.z, m.w}, ss[4] = {s.x, s.y, s.z, s.w}, aa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = aa[k] - mm[k];
      lp += -(d * d) / (2.0f * ss[k] * ss[k]) - logf(ss[k]) - kHalfLog2Pi;  // Normal.log_prob
    }
    __stcs(reinterpret_cast<float4*>(actions + off), a);
    if (mu_out && mu_out != mu) __stcs(reinterpret_cast<float4*>(mu_out + off), m);
    if (sigma_out) __stcs(reinterpret_cast<float4*>(sigma_out + off), s);
  }
  lp = lt::group_sum<float, 4>(lp);
  if (valid && sub == 0) logp[n] = lp;
}

__global__ void __launch_bounds__(256)
store_scalars_kernel(const float* __restrict__ rewards, const int64_t* __restrict__ dones_i64, const uint8_t* __restrict__ dones_u8,
                     const uint8_t* __restrict__ time_outs, const float* __restrict__ values, float gamma,
                     float* __restrict__ rewards_out, uint8_t* __restrict__ dones_out, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float r = rewards[n];
  if (time_outs) r = __fadd_rn(r, __fmul_rn(gamma, __fmul_rn(values[n], (float)time_outs[n])));  // ppo.py:163-165
  rewards_out[n] = r;
  dones_out[n] = dones_i64 ? (uint8_t)(dones_i64[n] != 0) : dones_u8[n];
}

__global__ void __launch_bounds__(256) copy4_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) __stcs(dst + i, __ldcs(src + i));
}
__global__ void __launch_bounds__(256) copy1_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

int copy_rows(const float* src, float* dst, int64_t n, cudaStream_t st) {
  if (!src || !dst || src == dst || n <= 0) return LT_OK;
  if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0 && (n & 3) == 0) {
    copy4_kernel<<<(unsigned)lt::ceil_div(n >> 2, 256), 256, 0, st>>>((const float4*)src, (float4*)dst, n >> 2);
  } else {
    copy1_kernel<<<(unsigned)lt::ceil_div(n, 256), 256, 0, st>>>(src, dst, n);
  }
  return lt::check_launch();
}

// One warp per output row; tensors are walked in order, each row copied with the widest aligned access available.
struct GatherParams {
  int num;
  const float* src[LT_GATHER_MAX];
  float* dst[LT_GATHER_MAX];
  int row_len[LT_GATHER_MAX];
  int vec4[LT_GATHER_MAX];
};

__global__ void __launch_bounds__(256)
gather_rows_kernel(const GatherParams p, const int64_t* __restrict__ indices, int64_t count) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < count; r += nwarps) {
    const int64_t s = __ldg(indices + r);
#pragma unroll 1
    for (int t = 0; t < p.num; ++t) {
      const int len = p.row_len[t];
      const float* src = p.src[t] + (size_t)s * len;
      float* dst = p.dst[t] + (size_t)r * len;
      if (p.vec4[t]) {
        for (int i = lane; i < (len >> 2); i += 32) __stcs(reinterpret_cast<float4*>(dst) + i, __ldg(reinterpret_cast<const float4*>(src) + i));
      } else if ((len & 1) == 0) {
        for (int i = lane; i < (len >> 1); i += 32) __stcs(reinterpret_cast<float2*>(dst) + i, __ldg(reinterpret_cast<const float2*>(src) + i));
      } else {
        for (int i = lane; i < len; i += 32) __stcs(dst + i, __ldg(src + i));
      }
    }
  }
}

__global__ void __launch_bounds__(256)
process_actions_kernel(const float* __restrict__ actions, float clip, float raw_scale, float scale, const float* __restrict__ offset,
                       float* __restrict__ raw, float* __restrict__ prev_raw, float* __restrict__ prev_prev_raw,
                       float* __restrict__ processed, float* __restrict__ prev_processed, float* __restrict__ prev_prev_processed,
                       int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float old_raw = raw[i];
  if (prev_prev_raw) prev_prev_raw[i] = prev_raw[i];
  prev_raw[i] = old_raw;
  if (processed) {
    const float old_p = processed[i];
    if (prev_prev_processed && prev_processed) prev_prev_processed[i] = prev_processed[i];
    if (prev_processed) prev_processed[i] = old_p;
  }
  float a = actions[i];
  if (clip > 0.f) a = fminf(fmaxf(a, -clip), clip);  // torch.clamp
  a = __fmul_rn(a, raw_scale);
  raw[i] = a;
  if (processed) processed[i] = __fadd_rn(__fmul_rn(a, scale), offset ? offset[i] : 0.f);
}

__global__ void counter_add_kernel(int64_t* c, int64_t inc) { *c += inc; }

}  // namespace

extern "C" int lt_counter_add(int64_t* counter, int64_t inc, void* stream) {
  if (!counter) return LT_ERR_INVALID_ARG;
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, inc);
  return lt::check_launch();
}

extern "C" int lt_process_actions(const float* actions, float clip, float raw_scale, float scale, const float* offset, float* raw,
                                  float* prev_raw, float* prev_prev_raw, float* processed, float* prev_processed,
                                  float* prev_prev_processed, int64_t count, void* stream) {
  if (!actions || !raw || !prev_raw || count <= 0) return LT_ERR_INVALID_ARG;
  process_actions_kernel<<<(unsigned)lt::ceil_div(count, 256), 256, 0, (cudaStream_t)stream>>>(
      actions, clip, raw_scale, scale, offset, raw, prev_raw, prev_prev_raw, processed, prev_processed, prev_prev_processed, count);
  return lt::check_launch();
}

extern "C" int lt_act_sample(const float* mu, const float* sigma, const float* eps, float* actions, float* logp, float* mu_out,
                             float* sigma_out, int N, int A, uint64_t seed, uint64_t offset, const int64_t* offset_base, void* stream) {
  if (!mu || !sigma || !actions || !logp || N <= 0 || A <= 0 || (A & 3)) return LT_ERR_INVALID_ARG;
  uintptr_t al = (uintptr_t)mu | (uintptr_t)sigma | (uintptr_t)actions | (uintptr_t)(eps ? eps : mu) |
                 (uintptr_t)(mu_out ? mu_out : mu) | (uintptr_t)(sigma_out ? sigma_out : mu);
  if (al & 15) return LT_ERR_INVALID_ARG;
  const int grid = (int)lt::ceil_div((int64_t)N * 4, 256);
  act_sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mu, sigma, eps, actions, logp, mu_out, sigma_out, N, A, seed, offset, offset_base);
  return lt::check_launch();
}

extern "C" int lt_store_step(const float* rewards, const int64_t* dones_i64, const uint8_t* dones_u8, const uint8_t* time_outs,
                             const float* values, float gamma, float* rewards_out, uint8_t* dones_out, const float* obs,
                             float* obs_out, int obs_dim, const float* critic_obs, float* critic_obs_out, int critic_obs_dim,
                             int N, void* stream) {
  if (!rewards || (!dones_i64 && !dones_u8) || !rewards_out || !dones_out || N <= 0) return LT_ERR_INVALID_ARG;
  if (time_outs && !values) return LT_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  store_scalars_kernel<<<(unsigned)lt::ceil_div(N, 256), 256, 0, st>>>(rewards, dones_i64, dones_u8, time_outs, values, gamma,
                                                                     rewards_out, dones_out, N);
  int rc = lt::check_launch();
  if (rc != LT_OK) return rc;
  rc = copy_rows(obs, obs_out, (int64_t)N * obs_dim, st);
  if (rc != LT_OK) return rc;
  return copy_rows(critic_obs, critic_obs_out, (int64_t)N * critic_obs_dim, st);
}

extern "C" int lt_gather_rows(const LtGatherArgs* a, const int64_t* indices, int64_t count, void* stream) {
  if (!a || !indices || count <= 0 || a->num_tensors <= 0 || a->num_tensors > LT_GATHER_MAX) return LT_ERR_INVALID_ARG;
  GatherParams p;
  p.num = a->num_tensors;
  for (int t = 0; t < p.num; ++t) {
    if (!a->src[t] || !a->dst[t] || a->row_len[t] <= 0) return LT_ERR_INVALID_ARG;
    p.src[t] = a->src[t];
    p.dst[t] = a->dst[t];
    p.row_len[t] = a->row_len[t];
    const bool al16 = ((((uintptr_t)a->src[t] | (uintptr_t)a->dst[t]) & 15) == 0) && (a->row_len[t] % 4 == 0);
    p.vec4[t] = al16 ? 1 : 0;
    if (!al16 && (a->row_len[t] % 2 == 0) && ((((uintptr_t)a->src[t] | (uintptr_t)a->dst[t]) & 7) != 0)) return LT_ERR_INVALID_ARG;
  }
  int64_t blocks = lt::ceil_div(count, 8);
  const int64_t cap = 16LL * lt::sm_count();
  if (blocks > cap) blocks = cap;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, indices, count);
  return lt::check_launch();
}
