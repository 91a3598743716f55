// K7: global-norm gradient clip + Adam(W) over ONE flat fp32 parameter buffer.
// Replaces reference loco_rl/loco_rl/algorithms/ppo.py:350-353 (nn.utils.clip_grad_norm_ + optim.Adam.step, ~6 foreach
// launches per parameter group) and the AdamW step of locotouch/distill/student.py:82,151.
//
// ONE launch, no host round trip (clip_adam_fused_kernel): every thread loads its (up to 4) float4 of the gradient ONCE and keeps
// them in registers; per-block partial sums of g^2 in fp64 go to the workspace; a grid-wide arrive / spin barrier on a device
// counter (the grid is bounded so that all blocks are co-resident); every block folds the partials in the same fixed order, derives
// the clip coefficient and the bias corrections, and updates p / m / v with 128-bit accesses from the registers it already holds.
// lr and the step counter are device scalars so the launch can be captured in a CUDA graph and replayed.  After a DDP
// all-reduce(sum) pass grad_scale = 1/world_size.  Parameter counts beyond the co-resident grid take the two-launch path
// (norm pass, then update pass).  Algorithmic traffic: 16 B read + 12 B written per parameter.
#include "lt_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;  // float4 per thread per tile

__global__ void __launch_bounds__(kThreads)
grad_sqnorm_kernel(const float* __restrict__ g, int64_t n, float grad_scale, double* __restrict__ partial, float* step_inout) {
  __shared__ double red[kThreads / 32];
  const int64_t n4 = n >> 2;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    const float a = v.x * grad_scale, b = v.y * grad_scale, c = v.z * grad_scale, d = v.w * grad_scale;
    acc += (double)(a * a + b * b) + (double)(c * c + d * d);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t j = n4 << 2; j < n; ++j) {
      const float a = g[j] * grad_scale;
      acc += (double)(a * a);
    }
    *step_inout += 1.0f;  // Adam's state["step"] += 1
  }
  acc = lt::block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float coef, float lr, float step_size,
                                         float sqrt_bc2, float omb1, float b2, float omb2, float eps, float wd) {
  g *= coef;
  if (wd != 0.f) p *= 1.0f - lr * wd;        // AdamW decoupled decay
  m = m + (g - m) * omb1;                    // exp_avg.lerp_(grad, 1 - beta1)   (1 - beta evaluated in double like Python)
  v = v * b2 + omb2 * g * g;                 // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / sqrt_bc2 + eps;  // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p = p - step_size * (m / denom);           // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(kThreads)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 const float* __restrict__ lr_ptr, const float* __restrict__ step_ptr, float max_norm, double b1d, double b2d,
                 float eps, float wd, float grad_scale, const double* __restrict__ partial, int num_partials,
                 float* grad_norm_out) {
  __shared__ double red[kThreads / 32];
  __shared__ float s_coef, s_step_size, s_sqrt_bc2;
  double acc = 0.0;
  for (int i = threadIdx.x; i < num_partials; i += kThreads) acc += partial[i];
  acc = lt::block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float total = (float)sqrt(acc);
    float c = 1.0f;
    if (max_norm > 0.f) c = fminf(max_norm / (total + 1e-6f), 1.0f);  // clip_grad_norm_: clamp(max_norm/(norm+1e-6), max=1)
    s_coef = c * grad_scale;
    if (blockIdx.x == 0 && grad_norm_out) *grad_norm_out = total;
    // torch.optim.Adam evaluates the bias corrections in Python doubles
    const double step = (double)*step_ptr;
    const double bc1 = 1.0 - pow(b1d, step), bc2 = 1.0 - pow(b2d, step);
    s_step_size = (float)((double)*lr_ptr / bc1);
    s_sqrt_bc2 = (float)sqrt(bc2);
  }
  __syncthreads();
  const float coef = s_coef, lr = *lr_ptr, step_size = s_step_size, sqrt_bc2 = s_sqrt_bc2;
  const float omb1 = (float)(1.0 - b1d), b2 = (float)b2d, omb2 = (float)(1.0 - b2d);
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, coef, lr, step_size, sqrt_bc2, omb1, b2, omb2, eps, wd);
    adam_one(pp.y, gg.y, mm.y, vv.y, coef, lr, step_size, sqrt_bc2, omb1, b2, omb2, eps, wd);
    adam_one(pp.z, gg.z, mm.z, vv.z, coef, lr, step_size, sqrt_bc2, omb1, b2, omb2, eps, wd);
    adam_one(pp.w, gg.w, mm.w, vv.w, coef, lr, step_size, sqrt_bc2, omb1, b2, omb2, eps, wd);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t j = n4 << 2; j < n; ++j) adam_one(p[j], g[j], m[j], v[j], coef, lr, step_size, sqrt_bc2, omb1, b2, omb2, eps, wd);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// One-launch variants: gradient (own buffer, or the sum of the W ranks' buffers read by peer loads) held in registers across a
// grid barrier.
constexpr int kFusedVec = 4;           // float4 per thread held in registers
constexpr int kFusedBlocksPerSm = 4;   // 4 x 256 threads per SM: co-resident for any register count <= 64

struct FusedWs {          // lives behind the 1024 fp64 partials
  unsigned int arrive, depart, pad[2];
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct PeerPtrs {
  const float* p[LT_MAX_PEERS];
};

template <bool kPeer>
__global__ void __launch_bounds__(kThreads, kFusedBlocksPerSm)
clip_adam_fused_kernel(float* __restrict__ p, const float* __restrict__ g, const __grid_constant__ PeerPtrs peer, int world, int tail, int64_t gather_slice4,
                       float* __restrict__ m, float* __restrict__ v, int64_t n, float* lr_ptr, float* step_ptr, float max_norm,
                       double b1d, double b2d, float eps, float wd, float grad_scale, float desired_kl, float kl_scale,
                       double* __restrict__ partial, FusedWs* ws, float* grad_norm_out, float* __restrict__ gsum_tail) {
  __shared__ double red[kThreads / 32];
  __shared__ float s_coef, s_step_size, s_sqrt_bc2, s_lr;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t i0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  float4 gg[kFusedVec];
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < kFusedVec; ++k) {
    const int64_t i = i0 + k * stride;
    gg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
      if constexpr (kPeer) {  // rank order 0..W-1 on every rank: bit-identical sums, replicas stay replicas
        float4 s;
        if (gather_slice4 > 0) {  // two-shot exchange, second half: slice q of the sum was reduced by rank q into ITS buffer
          const int64_t owner = min(i / gather_slice4, (int64_t)world - 1);
          asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(s.x), "=f"(s.y), "=f"(s.z), "=f"(s.w) : "l"(peer.p[owner] + 4 * i));
          gg[k] = s;
          const float a = s.x * grad_scale, b = s.y * grad_scale, c = s.z * grad_scale, d = s.w * grad_scale;
          acc += (double)(a * a + b * b) + (double)(c * c + d * d);
          continue;
        }
        asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(s.x), "=f"(s.y), "=f"(s.z), "=f"(s.w) : "l"(peer.p[0] + 4 * i));
        // the remote loads go out four at a time (one NVLink round trip per group instead of one per rank); the additions keep rank order
        for (int r0 = 1; r0 < world; r0 += 4) {
          float4 t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (r0 + u < world)
              asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t[u].x), "=f"(t[u].y), "=f"(t[u].z), "=f"(t[u].w) : "l"(peer.p[r0 + u] + 4 * i));
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (r0 + u < world) { s.x += t[u].x; s.y += t[u].y; s.z += t[u].z; s.w += t[u].w; }
        }
        gg[k] = s;
      } else {
        gg[k] = __ldcs(reinterpret_cast<const float4*>(g) + i);
      }
      const float a = gg[k].x * grad_scale, b = gg[k].y * grad_scale, c = gg[k].z * grad_scale, d = gg[k].w * grad_scale;
      acc += (double)(a * a + b * b) + (double)(c * c + d * d);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if constexpr (!kPeer) {
      for (int64_t j = n4 << 2; j < n; ++j) {
        const float a = g[j] * grad_scale;
        acc += (double)(a * a);
      }
    } else {
      for (int j = 0; j < tail; ++j) {  // statistics riding behind the gradients (KL mean): summed, not part of the norm
        float t = 0.f;
        for (int r = 0; r < world; ++r) {
          float x;
          asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x) : "l"(peer.p[r] + n + j));
          t += x;
        }
        if (gsum_tail) gsum_tail[j] = t;
        if (j == 0 && desired_kl > 0.f) {  // ppo.py:275-281 on the KL mean over all ranks, BEFORE the step that follows
          const float kl = t * kl_scale;
          float lr = *lr_ptr;
          if (kl > desired_kl * 2.0f)
            lr = fmaxf(1e-5f, lr / 1.5f);
          else if (kl < desired_kl / 2.0f && kl > 0.0f)
            lr = fminf(1e-2f, lr * 1.5f);
          *lr_ptr = lr;
        }
      }
    }
    *step_ptr += 1.0f;  // Adam's state["step"] += 1
  }
  acc = lt::block_sum(acc, red);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = acc;
    __threadfence();
    atomicAdd(&ws->arrive, 1u);
    while (ld_acquire_u32(&ws->arrive) < gridDim.x) {
    }
  }
  __syncthreads();
  double tot = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) tot += __ldcg(partial + i);
  tot = lt::block_sum(tot, red);
  if (threadIdx.x == 0) {
    const float total = (float)sqrt(tot);
    float c = 1.0f;
    if (max_norm > 0.f) c = fminf(max_norm / (total + 1e-6f), 1.0f);  // clip_grad_norm_: clamp(max_norm/(norm+1e-6), max=1)
    s_coef = c * grad_scale;
    if (blockIdx.x == 0 && grad_norm_out) *grad_norm_out = total;
    const float lr = __ldcg(lr_ptr);
    const double step = (double)__ldcg(step_ptr);
    const double bc1 = 1.0 - pow(b1d, step), bc2 = 1.0 - pow(b2d, step);  // torch.optim.Adam: Python doubles
    s_step_size = (float)((double)lr / bc1);
    s_sqrt_bc2 = (float)sqrt(bc2);
    s_lr = lr;
  }
  __syncthreads();
  const float coef = s_coef, lr = s_lr, step_size = s_step_size, sqrt_bc2 = s_sqrt_bc2;
  const float omb1 = (float)(1.0 - b1d), b2 = (float)b2d, omb2 = (float)(1.0 - b2d);
#pragma unroll
  for (int k = 0; k < kFusedVec; ++k) {
    const int64_t i = i0 + k * stride;
    if (i < n4) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      adam_one(pp.x, gg[k].x, mm.x, vv.x, coef, lr, step_size, sqrt_bc2, omb1, b2, omb2, eps, wd);
      adam_one(pp.y, gg[k].y, mm.y, vv.y, coef, lr, step_size, sqrt_bc2, omb1, b2, omb2, eps, wd);
      adam_one(pp.z, gg[k].z, mm.z, vv.z, coef, lr, step_size, sqrt_bc2, omb1, b2, omb2, eps, wd);
      adam_one(pp.w, gg[k].w, mm.w, vv.w, coef, lr, step_size, sqrt_bc2, omb1, b2, omb2, eps, wd);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if constexpr (!kPeer)
      for (int64_t j = n4 << 2; j < n; ++j) adam_one(p[j], g[j], m[j], v[j], coef, lr, step_size, sqrt_bc2, omb1, b2, omb2, eps, wd);
  }
  if (threadIdx.x == 0) {
    // the last block to leave re-arms the barrier: everybody has passed the spin by then
    if (atomicAdd(&ws->depart, 1u) == gridDim.x - 1) {
      ws->arrive = 0;
      ws->depart = 0;
      __threadfence();
    }
  }
}

// grid of the one-launch variant, or 0 when n does not fit the co-resident grid's registers
int fused_grid_for(int64_t n) {
  const int64_t n4 = n >> 2;
  const int64_t cap = (int64_t)kFusedBlocksPerSm * lt::sm_count();
  const int64_t want = lt::ceil_div(n4 > 0 ? n4 : 1, (int64_t)kThreads * kFusedVec);
  if (want > cap || want > 1024) return 0;
  // spread over at least one block per SM when there is enough work (more bytes in flight per SM-cycle)
  int64_t g = lt::ceil_div(n4 > 0 ? n4 : 1, (int64_t)kThreads * 2);
  if (g > cap) g = cap;
  if (g > 1024) g = 1024;
  if (g < want) g = want;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------------------------------
// K14: the gradient all-reduce of an env-sharded PPO step folded into the optimizer.  Every rank keeps its flat
// gradient buffer in NVLink-mapped symmetric memory; after a cross-GPU barrier each rank READS the W buffers directly over
// NVLink / NVSwitch (peer loads), adds them in rank order 0..W-1 -- the same order on every rank, so all ranks hold bit-identical
// sums and stay replicas -- and accumulates the squared norm of the sum in the same pass.  The clip + Adam kernel then runs on the
// local sum.  Replaces: NCCL all-reduce (2.75 MB, latency-bound: ~42 us on 2 GPUs) + the separate norm pass.
__device__ __forceinline__ float4 ld_peer4(const float* p) {  // system-scope load: the line may live in another GPU's memory
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kThreads)
peer_sum_sqnorm_kernel(const PeerPtrs peers, int world, int64_t n, int tail, float grad_scale, float* __restrict__ gsum,
                       double* __restrict__ partial, float* step_inout, float desired_kl, float kl_scale, float* lr_inout) {
  __shared__ double red[kThreads / 32];
  const int64_t n4 = n >> 2;  // n is a multiple of 4 (every parameter tensor starts 16-byte aligned in the flat buffer)
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
    float4 s = ld_peer4(peers.p[0] + 4 * i);
    for (int r = 1; r < world; ++r) {
      const float4 v = ld_peer4(peers.p[r] + 4 * i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(gsum)[i] = s;
    const float a = s.x * grad_scale, b = s.y * grad_scale, c = s.z * grad_scale, d = s.w * grad_scale;
    acc += (double)(a * a + b * b) + (double)(c * c + d * d);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int j = 0; j < tail; ++j) {  // statistics riding behind the gradients (KL mean): summed, not part of the norm
      float t = 0.f;
      for (int r = 0; r < world; ++r) {
        float v;
        asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(peers.p[r] + n + j));
        t += v;
      }
      gsum[n + j] = t;
      if (j == 0 && desired_kl > 0.f && lr_inout) {  // ppo.py:275-281 on the KL mean over all ranks, BEFORE the step that follows
        const float kl = t * kl_scale;
        float lr = *lr_inout;
        if (kl > desired_kl * 2.0f)
          lr = fmaxf(1e-5f, lr / 1.5f);
        else if (kl < desired_kl / 2.0f && kl > 0.0f)
          lr = fminf(1e-2f, lr * 1.5f);
        *lr_inout = lr;
      }
    }
    *step_inout += 1.0f;
  }
  acc = lt::block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// Two-shot exchange, first half (reduce-scatter by peer loads): rank q adds slice q of all W buffers in rank order and writes the sum
// over slice q of its OWN buffer (nobody else reads that slice in this phase).  After a cross-GPU barrier the gather mode of the
// fused kernel reads slice q from rank q: 2 (W - 1) / W buffers cross NVLink per rank instead of W - 1.
__global__ void __launch_bounds__(kThreads)
peer_reduce_slice_kernel(const PeerPtrs peers, int world, int rank, int64_t lo4, int64_t hi4) {
  for (int64_t i = lo4 + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < hi4; i += (int64_t)gridDim.x * kThreads) {
    float4 t[LT_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < LT_MAX_PEERS; ++r)
      if (r < world) t[r] = ld_peer4(peers.p[r] + 4 * i);   // all W loads in flight together
    float4 s = t[0];
#pragma unroll
    for (int r = 1; r < LT_MAX_PEERS; ++r)
      if (r < world) { s.x += t[r].x; s.y += t[r].y; s.z += t[r].z; s.w += t[r].w; }
    reinterpret_cast<float4*>(const_cast<float*>(peers.p[rank]))[i] = s;
  }
}

int grid_for(int64_t n) {
  const int64_t want = lt::ceil_div(lt::ceil_div(n, 4), (int64_t)kThreads * kVecPerThread);
  int64_t cap = 8LL * lt::sm_count();
  if (cap > 1024) cap = 1024;  // workspace holds 1024 partials
  const int64_t g = want < cap ? want : cap;
  return (int)(g > 0 ? g : 1);
}

}  // namespace

extern "C" int64_t lt_clip_adam_workspace_bytes(int64_t n) {
  (void)n;
  return 1024 * (int64_t)sizeof(double) + 64;  // one fp64 partial per block (grids are capped at 1024 blocks) + the barrier counters
}

extern "C" int lt_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr,
                            float* step_inout, float max_grad_norm, double beta1, double beta2, float eps, float weight_decay,
                            float grad_scale, float* grad_norm_out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !lr || !step_inout || !workspace || n <= 0) return LT_ERR_INVALID_ARG;
  const uintptr_t align = (uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq;
  if (align & 15) return LT_ERR_INVALID_ARG;
  const int grid = grid_for(n);
  if (workspace_bytes < (int64_t)grid * (int64_t)sizeof(double)) return LT_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = (double*)workspace;
  const int fgrid = fused_grid_for(n);
  if (fgrid > 0 && workspace_bytes >= 1024 * (int64_t)sizeof(double) + (int64_t)sizeof(FusedWs)) {
    FusedWs* fws = (FusedWs*)((char*)workspace + 1024 * sizeof(double));
    clip_adam_fused_kernel<false><<<fgrid, kThreads, 0, st>>>(params, grads, PeerPtrs{}, 1, 0, 0, exp_avg, exp_avg_sq, n, const_cast<float*>(lr), step_inout,
                                                              max_grad_norm, beta1, beta2, eps, weight_decay, grad_scale, 0.f, 1.f, partial, fws,
                                                              grad_norm_out, nullptr);
    return lt::check_launch();
  }
  grad_sqnorm_kernel<<<grid, kThreads, 0, st>>>(grads, n, grad_scale, partial, step_inout);
  int rc = lt::check_launch();
  if (rc != LT_OK) return rc;
  clip_adam_kernel<<<grid, kThreads, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, step_inout, max_grad_norm, beta1, beta2,
                                              eps, weight_decay, grad_scale, partial, grid, grad_norm_out);
  return lt::check_launch();
}

static int64_t peer_slice4(int64_t n, int world) { return lt::ceil_div(n >> 2, (int64_t)world); }

extern "C" int lt_peer_reduce_scatter(const float* const* peer_grads, int world, int rank, int64_t n, void* stream) {
  if (!peer_grads || world < 1 || world > LT_MAX_PEERS || rank < 0 || rank >= world || n <= 0 || (n & 3)) return LT_ERR_INVALID_ARG;
  PeerPtrs peers;
  for (int r = 0; r < LT_MAX_PEERS; ++r) {
    peers.p[r] = r < world ? peer_grads[r] : nullptr;
    if (r < world && (!peer_grads[r] || ((uintptr_t)peer_grads[r] & 15))) return LT_ERR_INVALID_ARG;
  }
  const int64_t slice4 = peer_slice4(n, world), n4 = n >> 2;
  const int64_t lo = slice4 * rank < n4 ? slice4 * rank : n4, hi = (rank == world - 1) ? n4 : (slice4 * (rank + 1) < n4 ? slice4 * (rank + 1) : n4);
  if (hi <= lo) return LT_OK;
  int64_t blocks = lt::ceil_div(hi - lo, (int64_t)kThreads);
  const int64_t cap = 4LL * lt::sm_count();
  if (blocks > cap) blocks = cap;
  peer_reduce_slice_kernel<<<(int)blocks, kThreads, 0, (cudaStream_t)stream>>>(peers, world, rank, lo, hi);
  return lt::check_launch();
}

static int peer_clip_adam_impl(bool gather, float* params, const float* const* peer_grads, int world, float* grad_sum, int tail, float* exp_avg,
                               float* exp_avg_sq, int64_t n, float* lr, float* step_inout, float max_grad_norm, double beta1,
                               double beta2, float eps, float weight_decay, float grad_scale, float desired_kl, float kl_scale,
                               float* grad_norm_out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!params || !peer_grads || !grad_sum || !exp_avg || !exp_avg_sq || !lr || !step_inout || !workspace || n <= 0 || (n & 3)) return LT_ERR_INVALID_ARG;
  if (world < 1 || world > LT_MAX_PEERS || tail < 0 || tail > 16) return LT_ERR_INVALID_ARG;
  PeerPtrs peers;
  uintptr_t align = (uintptr_t)params | (uintptr_t)grad_sum | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq;
  for (int r = 0; r < LT_MAX_PEERS; ++r) {
    peers.p[r] = r < world ? peer_grads[r] : nullptr;
    if (r < world) {
      if (!peer_grads[r]) return LT_ERR_INVALID_ARG;
      align |= (uintptr_t)peer_grads[r];
    }
  }
  if (align & 15) return LT_ERR_INVALID_ARG;
  const int grid = grid_for(n);
  if (workspace_bytes < (int64_t)grid * (int64_t)sizeof(double)) return LT_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = (double*)workspace;
  const int fgrid = fused_grid_for(n);
  if (fgrid > 0 && workspace_bytes >= 1024 * (int64_t)sizeof(double) + (int64_t)sizeof(FusedWs)) {
    // one launch: peer sums stay in registers across the grid barrier (grad_sum only receives the summed tail statistics)
    FusedWs* fws = (FusedWs*)((char*)workspace + 1024 * sizeof(double));
    clip_adam_fused_kernel<true><<<fgrid, kThreads, 0, st>>>(params, nullptr, peers, world, tail, gather ? peer_slice4(n, world) : 0, exp_avg, exp_avg_sq, n, lr, step_inout, max_grad_norm,
                                                             beta1, beta2, eps, weight_decay, grad_scale, desired_kl, kl_scale, partial, fws,
                                                             grad_norm_out, grad_sum + n);
    return lt::check_launch();
  }
  if (gather) return LT_ERR_UNSUPPORTED;  // the gather mode exists in the one-launch kernel only (parameter counts beyond it: one-shot)
  peer_sum_sqnorm_kernel<<<grid, kThreads, 0, st>>>(peers, world, n, tail, grad_scale, grad_sum, partial, step_inout, desired_kl, kl_scale, lr);
  int rc = lt::check_launch();
  if (rc != LT_OK) return rc;
  clip_adam_kernel<<<grid, kThreads, 0, st>>>(params, grad_sum, exp_avg, exp_avg_sq, n, lr, step_inout, max_grad_norm, beta1, beta2, eps,
                                              weight_decay, grad_scale, partial, grid, grad_norm_out);
  return lt::check_launch();
}

extern "C" int lt_peer_sum_clip_adam(float* params, const float* const* peer_grads, int world, float* grad_sum, int tail, float* exp_avg,
                                     float* exp_avg_sq, int64_t n, float* lr, float* step_inout, float max_grad_norm, double beta1,
                                     double beta2, float eps, float weight_decay, float grad_scale, float desired_kl, float kl_scale,
                                     float* grad_norm_out, void* workspace, int64_t workspace_bytes, void* stream) {
  return peer_clip_adam_impl(false, params, peer_grads, world, grad_sum, tail, exp_avg, exp_avg_sq, n, lr, step_inout, max_grad_norm, beta1, beta2, eps,
                             weight_decay, grad_scale, desired_kl, kl_scale, grad_norm_out, workspace, workspace_bytes, stream);
}

extern "C" int lt_peer_gather_clip_adam(float* params, const float* const* peer_grads, int world, float* grad_sum, int tail, float* exp_avg,
                                        float* exp_avg_sq, int64_t n, float* lr, float* step_inout, float max_grad_norm, double beta1,
                                        double beta2, float eps, float weight_decay, float grad_scale, float desired_kl, float kl_scale,
                                        float* grad_norm_out, void* workspace, int64_t workspace_bytes, void* stream) {
  return peer_clip_adam_impl(true, params, peer_grads, world, grad_sum, tail, exp_avg, exp_avg_sq, n, lr, step_inout, max_grad_norm, beta1, beta2, eps,
                             weight_decay, grad_scale, desired_kl, kl_scale, grad_norm_out, workspace, workspace_bytes, stream);
}
