// K9: fused activation backward + bias gradient for the actor-critic / student MLPs.
// Replaces, per Linear+ELU layer of the reference networks (loco_rl/loco_rl/modules/actor_critic.py:33-56, models/mlp.py:4-25),
// the two autograd kernels `elu_backward` (grad * (h > 0 ? 1 : h + alpha)) and the bias-gradient column reduction
// (sum over the batch), which together read the [B, n] gradient three times; here it is read once, written once, and the
// column sums fall out of the same pass.  The GEMMs themselves (dgrad / wgrad) stay cuBLAS.
// Column sums: each block folds its rows in a fixed order in shared memory and adds its partial to one of kSub float64
// accumulators per column (L2 atomics; the fp64 sum is order-independent to ~1e-16, i.e. the fp32 result is reproducible --
// spreading the blocks over kSub accumulator rows keeps the same-address atomic chains short: with one row, 592 serial atomics
// per column cost ~25 us); the last block to finish folds the kSub rows, rounds to fp32 and clears them.
// Traffic per element: 8 B read (+0 for the last, activation-free layer: 4 B) + 4 B written.
#include "lt_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kSub = 16;  // accumulator rows

struct BwdWs {
  unsigned int counter;
  unsigned int pad[3];
  double acc[1];  // [kSub][n] column accumulators, zero between calls
};

template <bool VEC4>
__global__ void __launch_bounds__(kThreads)
bias_act_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ act_out, float* __restrict__ grad_pre,
                    float* __restrict__ bias_grad, int B, int n, float alpha, int rows_per_block, BwdWs* ws) {
  extern __shared__ float s_acc[];  // [row_lanes][cols * (VEC4 ? 4 : 1)]
  __shared__ bool is_last;
  const int cols = VEC4 ? (n >> 2) : n;              // column groups
  const int col_threads = cols < kThreads ? cols : kThreads;
  const int row_lanes = kThreads / col_threads;
  const int cg0 = threadIdx.x % col_threads, rl = threadIdx.x / col_threads;
  const int row0 = blockIdx.x * rows_per_block;
  const int row1 = min(B, row0 + rows_per_block);
  const bool active = rl < row_lanes;
  for (int cg = cg0; cg < cols; cg += col_threads) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (active) {
      if constexpr (VEC4) {
        // 4 independent rows per trip: 8 x 16-byte loads in flight per thread
        int r = row0 + rl;
        for (; r + 3 * row_lanes < row1; r += 4 * row_lanes) {
          float4 g[4], h[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const size_t i = (size_t)(r + u * row_lanes) * cols + cg;
            g[u] = __ldcs(reinterpret_cast<const float4*>(grad_out) + i);
            h[u] = act_out ? __ldcs(reinterpret_cast<const float4*>(act_out) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const size_t i = (size_t)(r + u * row_lanes) * cols + cg;
            if (act_out) {
              g[u].x *= h[u].x > 0.f ? 1.f : h[u].x + alpha;
              g[u].y *= h[u].y > 0.f ? 1.f : h[u].y + alpha;
              g[u].z *= h[u].z > 0.f ? 1.f : h[u].z + alpha;
              g[u].w *= h[u].w > 0.f ? 1.f : h[u].w + alpha;
            }
            if (grad_pre && (act_out || grad_pre != grad_out)) reinterpret_cast<float4*>(grad_pre)[i] = g[u];
            a0 += g[u].x; a1 += g[u].y; a2 += g[u].z; a3 += g[u].w;
          }
        }
        for (; r < row1; r += row_lanes) {
          const size_t i = (size_t)r * cols + cg;
          float4 g = __ldcs(reinterpret_cast<const float4*>(grad_out) + i);
          if (act_out) {
            const float4 h = __ldcs(reinterpret_cast<const float4*>(act_out) + i);
            g.x *= h.x > 0.f ? 1.f : h.x + alpha;
            g.y *= h.y > 0.f ? 1.f : h.y + alpha;
            g.z *= h.z > 0.f ? 1.f : h.z + alpha;
            g.w *= h.w > 0.f ? 1.f : h.w + alpha;
          }
          if (grad_pre && (act_out || grad_pre != grad_out)) reinterpret_cast<float4*>(grad_pre)[i] = g;
          a0 += g.x; a1 += g.y; a2 += g.z; a3 += g.w;
        }
      } else {
        for (int r = row0 + rl; r < row1; r += row_lanes) {
          const size_t i = (size_t)r * n + cg;
          float g = grad_out[i];
          if (act_out) {
            const float h = act_out[i];
            g *= h > 0.f ? 1.f : h + alpha;
          }
          if (grad_pre && (act_out || grad_pre != grad_out)) grad_pre[i] = g;
          a0 += g;
        }
      }
    }
    // fold the row lanes (fixed order) and publish this block's partial for the column group
    if constexpr (VEC4) {
      float4* sa = reinterpret_cast<float4*>(s_acc);
      if (active) sa[rl * col_threads + cg0] = make_float4(a0, a1, a2, a3);
      __syncthreads();
      if (rl == 0) {
        float4 t = sa[cg0];
        for (int k = 1; k < row_lanes; ++k) {
          const float4 u = sa[k * col_threads + cg0];
          t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        double* acc = ws->acc + (size_t)(blockIdx.x % kSub) * n + 4 * cg;
        atomicAdd(acc, (double)t.x); atomicAdd(acc + 1, (double)t.y); atomicAdd(acc + 2, (double)t.z); atomicAdd(acc + 3, (double)t.w);
      }
      __syncthreads();
    } else {
      if (active) s_acc[rl * col_threads + cg0] = a0;
      __syncthreads();
      if (rl == 0) {
        float t = s_acc[cg0];
        for (int k = 1; k < row_lanes; ++k) t += s_acc[k * col_threads + cg0];
        atomicAdd(ws->acc + (size_t)(blockIdx.x % kSub) * n + cg, (double)t);
      }
      __syncthreads();
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&ws->counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int j = threadIdx.x; j < n; j += kThreads) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kSub; ++k) {
      t += __ldcg(ws->acc + (size_t)k * n + j);
      ws->acc[(size_t)k * n + j] = 0.0;
    }
    bias_grad[j] = (float)t;
  }
  if (threadIdx.x == 0) ws->counter = 0;
}

int blocks_for(int B) {
  int blocks = 4 * lt::sm_count();
  if (blocks > B) blocks = B;
  if (blocks > 1024) blocks = 1024;
  return blocks < 1 ? 1 : blocks;
}

}  // namespace

extern "C" int64_t lt_bias_act_bwd_workspace_bytes(int B, int n) {
  (void)B;
  return 16 + (int64_t)kSub * (int64_t)(n > 0 ? n : 1) * (int64_t)sizeof(double);
}

extern "C" int lt_bias_act_bwd(const float* grad_out, const float* act_out, float* grad_pre, float* bias_grad, int B, int n, float alpha,
                               void* workspace, int64_t workspace_bytes, void* stream) {
  if (!grad_out || !bias_grad || !workspace || B <= 0 || n <= 0) return LT_ERR_INVALID_ARG;
  if (workspace_bytes < lt_bias_act_bwd_workspace_bytes(B, n)) return LT_ERR_WORKSPACE;
  const int blocks = blocks_for(B);
  const int rows_per_block = (int)lt::ceil_div(B, blocks);
  const int grid = (int)lt::ceil_div(B, rows_per_block);
  const uintptr_t al = (uintptr_t)grad_out | (uintptr_t)(act_out ? act_out : grad_out) | (uintptr_t)(grad_pre ? grad_pre : grad_out);
  const bool vec4 = (n % 4 == 0) && (al & 15) == 0;
  const int cols = vec4 ? n / 4 : n;
  const int col_threads = cols < kThreads ? cols : kThreads;
  const int row_lanes = kThreads / col_threads;
  const size_t smem = (size_t)row_lanes * col_threads * (vec4 ? 16 : 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec4)
    bias_act_bwd_kernel<true><<<grid, kThreads, smem, st>>>(grad_out, act_out, grad_pre, bias_grad, B, n, alpha, rows_per_block, (BwdWs*)workspace);
  else
    bias_act_bwd_kernel<false><<<grid, kThreads, smem, st>>>(grad_out, act_out, grad_pre, bias_grad, B, n, alpha, rows_per_block, (BwdWs*)workspace);
  return lt::check_launch();
}
