// K4: GAE returns + advantage normalisation.
// Replaces RolloutStorage.compute_returns (reference loco_rl/loco_rl/storage/rollout_storage.py:152-174).
//
// Layout: rewards / values / dones are [T, N] (time major, env contiguous), exactly RolloutStorage's buffers with the
// trailing singleton dropped.  One thread owns one env: every load of a warp is one coalesced 128 B line, all
// 3*T loads of a thread are issued before the scan starts (they do not depend on it), and the reverse recurrence is
// then evaluated in registers in the reference's operation order (no FMA contraction) so that returns are bit-exact.
// Algorithmic traffic: 9 B read + 8 B written per (t, env) + 4 B/env  (SURVEY.md 8d).
//
// lt_gae with normalisation at T = 24 is ONE launch (gae_fused_kernel): blocks of one warp (4096 envs = 128 blocks, one per SM
// instead of 32 blocks of 128 threads on a 148-SM part), the 24 advantages of an env stay in registers across a grid-wide
// arrive / spin barrier on a device counter, every block folds the block partials of sum / sum of squares in the same fixed order
// (so all blocks normalise with bit-identical mean / std) and writes its normalised advantages: the advantages are written once
// instead of written, re-read and re-written by a second launch.  The grid is bounded (<= 8 blocks per SM) so that all blocks are
// co-resident; larger env counts and other T take the two-launch path (scan + normalise), which is also what an env-sharded job
// uses (the statistics are summed over ranks between the two).
#include "lt_common.cuh"

namespace {

constexpr int kThreads = 128;

struct GaeWs {
  unsigned int counter;
  unsigned int pad[3];
  double partial[1];  // [2 * blocks]
};

__device__ __forceinline__ void gae_step(float r, float v, float d, float nv, float gamma, float lam, float& adv, float& ret,
                                         float& out_adv) {
  // rollout_storage.py:161-167, evaluated left to right in fp32
  const float nt = __fsub_rn(1.0f, d);
  const float ng = __fmul_rn(nt, gamma);
  const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(ng, nv)), v);
  adv = __fadd_rn(delta, __fmul_rn(__fmul_rn(ng, lam), adv));
  ret = __fadd_rn(adv, v);
  out_adv = __fsub_rn(ret, v);  // rollout_storage.py:170
}

template <int TT>
__global__ void __launch_bounds__(kThreads)
gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                const float* __restrict__ last_values, float* __restrict__ returns, float* __restrict__ advantages,
                int T, int N, float gamma, float lam, GaeWs* ws, double* __restrict__ stats) {
  __shared__ double red[2 * kThreads / 32];
  __shared__ bool is_last;
  const int n = blockIdx.x * kThreads + threadIdx.x;
  double s = 0.0, ss = 0.0;
  if (n < N) {
    if constexpr (TT > 0) {
      float r[TT], v[TT + 1], d[TT];
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        r[t] = __ldcs(rewards + (size_t)t * N + n);
        v[t] = __ldcs(values + (size_t)t * N + n);
        d[t] = (float)__ldcs(dones + (size_t)t * N + n);
      }
      v[TT] = __ldcs(last_values + n);
      float adv = 0.0f;
#pragma unroll
      for (int t = TT - 1; t >= 0; --t) {
        float ret, a;
        gae_step(r[t], v[t], d[t], v[t + 1], gamma, lam, adv, ret, a);
        __stcs(returns + (size_t)t * N + n, ret);
        __stcs(advantages + (size_t)t * N + n, a);
        s += (double)a;
        ss += (double)a * (double)a;
      }
    } else {
      float adv = 0.0f;
      float nv = last_values[n];
      for (int t = T - 1; t >= 0; --t) {
        const float r = rewards[(size_t)t * N + n], v = values[(size_t)t * N + n];
        const float d = (float)dones[(size_t)t * N + n];
        float ret, a;
        gae_step(r, v, d, nv, gamma, lam, adv, ret, a);
        returns[(size_t)t * N + n] = ret;
        advantages[(size_t)t * N + n] = a;
        s += (double)a;
        ss += (double)a * (double)a;
        nv = v;
      }
    }
  }
  // deterministic block partials -> last block reduces them in a fixed order
  s = lt::block_sum(s, red);
  ss = lt::block_sum(ss, red + kThreads / 32);
  if (threadIdx.x == 0) {
    ws->partial[2 * blockIdx.x] = s;
    ws->partial[2 * blockIdx.x + 1] = ss;
    __threadfence();
    is_last = atomicAdd(&ws->counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) {
      a += __ldcg(&ws->partial[2 * i]);
      b += __ldcg(&ws->partial[2 * i + 1]);
    }
    a = lt::block_sum(a, red);
    b = lt::block_sum(b, red + kThreads / 32);
    if (threadIdx.x == 0) {
      stats[0] = a;
      stats[1] = b;
      stats[2] = (double)T * (double)N;
      stats[3] = 0.0;
      ws->counter = 0;  // self-cleaning
    }
  }
}

// ------------------------------------------------------------------------------------------------- one-launch variant
constexpr int kFusedThreads = 32;

struct GaeFusedWs {
  unsigned int arrive, depart;
  unsigned int pad[2];
  double partial[1];  // [2 * blocks]
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int TT>
__global__ void __launch_bounds__(kFusedThreads)
gae_fused_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                 const float* __restrict__ last_values, float* __restrict__ returns, float* __restrict__ advantages,
                 int N, float gamma, float lam, GaeFusedWs* ws, double* __restrict__ stats) {
  const int n = blockIdx.x * kFusedThreads + threadIdx.x;
  const bool valid = n < N;
  float a[TT];
  double s = 0.0, ss = 0.0;
  {
    float r[TT], v[TT + 1], d[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      r[t] = valid ? __ldcs(rewards + (size_t)t * N + n) : 0.f;
      v[t] = valid ? __ldcs(values + (size_t)t * N + n) : 0.f;
      d[t] = valid ? (float)__ldcs(dones + (size_t)t * N + n) : 0.f;
    }
    v[TT] = valid ? __ldcs(last_values + n) : 0.f;
    float adv = 0.0f;
#pragma unroll
    for (int t = TT - 1; t >= 0; --t) {
      float ret;
      gae_step(r[t], v[t], d[t], v[t + 1], gamma, lam, adv, ret, a[t]);
      if (valid) {
        __stcs(returns + (size_t)t * N + n, ret);
        s += (double)a[t];
        ss += (double)a[t] * (double)a[t];
      }
    }
  }
  // same association as the two-launch path for one warp: lane partial (t descending), then the shuffle tree
  s = lt::warp_sum(s);
  ss = lt::warp_sum(ss);
  if (threadIdx.x == 0) {
    ws->partial[2 * blockIdx.x] = s;
    ws->partial[2 * blockIdx.x + 1] = ss;
    __threadfence();
    atomicAdd(&ws->arrive, 1u);
    while (ld_acquire_u32(&ws->arrive) < gridDim.x) {
    }
  }
  __syncwarp();
  // every block folds all partials in the same order -> identical statistics everywhere
  double fa = 0.0, fb = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kFusedThreads) {
    fa += __ldcg(&ws->partial[2 * i]);
    fb += __ldcg(&ws->partial[2 * i + 1]);
  }
  fa = lt::warp_sum(fa);
  fb = lt::warp_sum(fb);
  const double cnt = (double)TT * (double)N;
  const double mean = fa / cnt;
  double var = (fb - fa * mean) / (cnt - 1.0);  // unbiased, rollout_storage.py:174 (.std())
  var = var > 0.0 ? var : 0.0;
  const float mean_f = (float)mean;
  const float denom = __fadd_rn((float)sqrt(var), 1e-8f);
  if (valid) {
#pragma unroll
    for (int t = 0; t < TT; ++t) __stcs(advantages + (size_t)t * N + n, __fdiv_rn(__fsub_rn(a[t], mean_f), denom));
  }
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      stats[0] = fa;
      stats[1] = fb;
      stats[2] = cnt;
      stats[3] = 0.0;
    }
    // the last block to leave re-arms the barrier: everybody has passed the spin by then
    if (atomicAdd(&ws->depart, 1u) == gridDim.x - 1) {
      ws->arrive = 0;
      ws->depart = 0;
      __threadfence();
    }
  }
}

__global__ void __launch_bounds__(256)
adv_normalize_kernel(float* __restrict__ adv, int64_t count, const double* __restrict__ stats) {
  const double s = stats[0], ss = stats[1], cnt = stats[2];
  const double mean = s / cnt;
  double var = (ss - s * mean) / (cnt - 1.0);  // unbiased, rollout_storage.py:174 (.std())
  var = var > 0.0 ? var : 0.0;
  const float mean_f = (float)mean;
  const float denom = __fadd_rn((float)sqrt(var), 1e-8f);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = count >> 2;
  if (i < n4) {
    float4 a = reinterpret_cast<float4*>(adv)[i];
    a.x = __fdiv_rn(__fsub_rn(a.x, mean_f), denom);
    a.y = __fdiv_rn(__fsub_rn(a.y, mean_f), denom);
    a.z = __fdiv_rn(__fsub_rn(a.z, mean_f), denom);
    a.w = __fdiv_rn(__fsub_rn(a.w, mean_f), denom);
    reinterpret_cast<float4*>(adv)[i] = a;
  }
  if (i == 0) {
    for (int64_t j = n4 << 2; j < count; ++j) adv[j] = __fdiv_rn(__fsub_rn(adv[j], mean_f), denom);
  }
}

}  // namespace

extern "C" int64_t lt_gae_workspace_bytes(int T, int N) {
  (void)T;
  const int64_t blocks = lt::ceil_div(N > 0 ? N : 1, kFusedThreads);  // the one-launch variant has the most blocks
  return 16 + 2 * blocks * (int64_t)sizeof(double) + 4 * (int64_t)sizeof(double);
}

extern "C" int lt_gae_scan(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                           float* returns, float* advantages, int T, int N, float gamma, float lam, double* stats,
                           void* workspace, int64_t workspace_bytes, void* stream) {
  if (!rewards || !values || !dones || !last_values || !returns || !advantages || !stats || !workspace || T <= 0 || N <= 0)
    return LT_ERR_INVALID_ARG;
  const int blocks = (int)lt::ceil_div(N, kThreads);
  if (workspace_bytes < 16 + 2 * (int64_t)blocks * (int64_t)sizeof(double)) return LT_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  GaeWs* ws = (GaeWs*)workspace;
  if (T == 24)
    gae_scan_kernel<24><<<blocks, kThreads, 0, st>>>(rewards, values, dones, last_values, returns, advantages, T, N, gamma, lam, ws, stats);
  else
    gae_scan_kernel<0><<<blocks, kThreads, 0, st>>>(rewards, values, dones, last_values, returns, advantages, T, N, gamma, lam, ws, stats);
  return lt::check_launch();
}

extern "C" int lt_adv_normalize(float* advantages, int64_t count, const double* stats, void* stream) {
  if (!advantages || !stats || count <= 1) return LT_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(advantages) & 15) != 0) return LT_ERR_INVALID_ARG;
  const int64_t n4 = (count + 3) >> 2;
  const int blocks = (int)lt::ceil_div(n4 > 0 ? n4 : 1, 256);
  adv_normalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(advantages, count, stats);
  return lt::check_launch();
}

extern "C" int lt_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                      float* returns, float* advantages, int T, int N, float gamma, float lam, int normalize,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  if (!workspace || workspace_bytes < lt_gae_workspace_bytes(T, N)) return LT_ERR_WORKSPACE;
  // stats live at the tail of the workspace
  double* stats = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + workspace_bytes - 4 * sizeof(double));
  if ((reinterpret_cast<uintptr_t>(stats) & 7) != 0) return LT_ERR_INVALID_ARG;
  if (!rewards || !values || !dones || !last_values || !returns || !advantages || T <= 0 || N <= 0) return LT_ERR_INVALID_ARG;
  const int64_t fused_blocks = lt::ceil_div(N, kFusedThreads);
  if (normalize && T == 24 && (int64_t)T * N > 1 && fused_blocks <= 8LL * lt::sm_count() &&
      workspace_bytes >= 16 + 2 * fused_blocks * (int64_t)sizeof(double) + 4 * (int64_t)sizeof(double)) {
    gae_fused_kernel<24><<<(int)fused_blocks, kFusedThreads, 0, (cudaStream_t)stream>>>(rewards, values, dones, last_values, returns, advantages, N, gamma,
                                                                                          lam, (GaeFusedWs*)workspace, stats);
    return lt::check_launch();
  }
  int rc = lt_gae_scan(rewards, values, dones, last_values, returns, advantages, T, N, gamma, lam, stats, workspace,
                       workspace_bytes - 4 * (int64_t)sizeof(double), stream);
  if (rc != LT_OK || !normalize) return rc;
  return lt_adv_normalize(advantages, (int64_t)T * N, stats, stream);
}
