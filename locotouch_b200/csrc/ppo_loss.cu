// K6: fused PPO loss, forward + analytic backward in ONE pass over the mini-batch.
// Replaces reference loco_rl/loco_rl/algorithms/ppo.py:252-302 (log-prob, entropy, KL(old||new), adaptive learning
// rate, clipped surrogate, clipped value loss, total loss) and the autograd backward of those expressions down to the
// network outputs (mu [B,A], value [B]) and the std parameter (sigma [A]).
//
// Mapping: 4 lanes per sample; lane `sub` owns the float4 chunks sub, sub+4, ... of the [A] rows, so every 128-bit load
// of a warp covers 8 consecutive rows = one contiguous span (A=12: 384 B).  Per-sample sums (log-prob, KL) are reduced
// with two shuffles inside the lane group.  Batch sums are reduced deterministically: warp shuffle -> shared -> one
// partial row per block in the workspace -> the last block to finish folds the rows in index order, writes the means,
// dL/dsigma and (adaptive schedule) the new learning rate.  No atomics on floating point, no host synchronisation.
// Algorithmic traffic per sample: 4 x A x 4 B + 20 B read, A x 4 B + 4 B written (A=12: 264 B).
#include "ppo_loss_common.cuh"

namespace {

using namespace lt_ppo;

// kChunks = float4 chunks per lane (A <= 16 * kChunks): the row registers are sized for the actual action dimension -- with the
// arrays dimensioned for A = 64 the kernel needed 120 registers and two blocks per SM, i.e. two waves for 384 blocks.
template <int kChunks>
__global__ void __launch_bounds__(kThreads) ppo_loss_kernel(const Params p) {
  __shared__ float s_sigma[kMaxA], s_inv_var[kMaxA], s_log_sigma[kMaxA];
  __shared__ float s_red[kThreads / 32][3 + kMaxA];
  const int A = p.A, chunks = A >> 2;
  for (int j = threadIdx.x; j < A; j += kThreads) {
    const float s = p.sigma[j];
    s_sigma[j] = s;
    s_inv_var[j] = 1.0f / (s * s);
    s_log_sigma[j] = logf(s);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 3;
  const float inv_b = 1.0f / (float)p.B;
  float acc_surr = 0.f, acc_vloss = 0.f, acc_kl = 0.f;
  float acc_dsig[kChunks][4];
#pragma unroll
  for (int c = 0; c < kChunks; ++c) acc_dsig[c][0] = acc_dsig[c][1] = acc_dsig[c][2] = acc_dsig[c][3] = 0.f;

  const int groups_per_block = kThreads / 4;
  // block-uniform trip count: every lane of a warp takes part in the lane-group shuffles below
  for (int base = blockIdx.x * groups_per_block; base < p.B; base += gridDim.x * groups_per_block) {
    const int b = base + (threadIdx.x >> 2);
    const bool valid = b < p.B;
    // ---- load: up to kChunks float4 of each [A] row + 5 scalars (same address in the 4 lanes: one transaction)
    float4 mu[kChunks], ac[kChunks], omu[kChunks], osg[kChunks];
    const size_t row = (size_t)b * A;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int ch = sub + 4 * c;
      mu[c] = ac[c] = omu[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      osg[c] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (valid && ch < chunks) {
        mu[c] = __ldg(reinterpret_cast<const float4*>(p.mu + row) + ch);
        ac[c] = __ldcs(reinterpret_cast<const float4*>(p.actions + row) + ch);
        omu[c] = __ldcs(reinterpret_cast<const float4*>(p.old_mu + row) + ch);
        osg[c] = __ldcs(reinterpret_cast<const float4*>(p.old_sigma + row) + ch);
      }
    }
    float old_logp = 0.f, adv = 0.f, val = 0.f, old_val = 0.f, ret = 0.f;
    if (valid) {
      old_logp = __ldcs(p.old_logp + b);
      adv = __ldcs(p.adv + b);
      val = __ldg(p.value + b);
      old_val = __ldcs(p.old_values + b);
      ret = __ldcs(p.returns + b);
    }

    // ---- per-sample log-prob and KL (partial over this lane's elements)
    float logp = 0.f, kl = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int ch = sub + 4 * c;
      if (ch < chunks) {
        const float m[4] = {mu[c].x, mu[c].y, mu[c].z, mu[c].w}, a[4] = {ac[c].x, ac[c].y, ac[c].z, ac[c].w};
        const float om[4] = {omu[c].x, omu[c].y, omu[c].z, omu[c].w}, os[4] = {osg[c].x, osg[c].y, osg[c].z, osg[c].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * ch + e;
          const float d = a[e] - m[e];
          const float half_iv = 0.5f * s_inv_var[j];
          logp += -(d * d) * half_iv - s_log_sigma[j] - kHalfLog2Pi;  // Normal.log_prob
          const float dm = om[e] - m[e];
          kl += logf(s_sigma[j] / os[e] + 1.0e-5f) + (os[e] * os[e] + dm * dm) * half_iv - 0.5f;  // ppo.py:266-272
        }
      }
    }
    logp = lt::group_sum<float, 4>(logp);
    kl = lt::group_sum<float, 4>(kl);

    // ---- surrogate (ppo.py:284-289) and its derivative w.r.t. log-prob
    const float ratio = expf(logp - old_logp);
    const float s_un = -adv * ratio;
    const float s_cl = -adv * fminf(fmaxf(ratio, p.clip_lo), p.clip_hi);
    const float surr = fmaxf(s_un, s_cl);
    const float g_logp = (s_un >= s_cl ? s_un : 0.f) * inv_b * p.grad_scale;  // d/dlogp(-A*ratio) = -A*ratio

    // ---- value loss (ppo.py:292-300) and its derivative w.r.t. value
    float vloss, g_val;
    const float e1 = val - ret;
    if (p.use_clipped_value) {
      const float d = val - old_val;
      const float dc = fminf(fmaxf(d, -p.clip), p.clip);
      const float e2 = (old_val + dc) - ret;
      const float v1 = e1 * e1, v2 = e2 * e2;
      vloss = fmaxf(v1, v2);
      const float in_rng = (d >= -p.clip && d <= p.clip) ? 1.f : 0.f;
      g_val = v1 > v2 ? 2.f * e1 : (v1 < v2 ? 2.f * e2 * in_rng : e1 + e2 * in_rng);
    } else {
      vloss = e1 * e1;
      g_val = 2.f * e1;
    }
    g_val *= p.vcoef * inv_b * p.grad_scale;

    if (sub == 0 && valid) {
      acc_surr += surr;
      acc_vloss += vloss;
      acc_kl += kl;
      __stcs(p.grad_value + b, g_val);
    }

    // ---- gradients w.r.t. mu (stored) and sigma (accumulated)
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int ch = sub + 4 * c;
      if (valid && ch < chunks) {
        const float m[4] = {mu[c].x, mu[c].y, mu[c].z, mu[c].w}, a[4] = {ac[c].x, ac[c].y, ac[c].z, ac[c].w};
        float g[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * ch + e;
          const float d = a[e] - m[e];
          const float iv = s_inv_var[j];
          g[e] = g_logp * d * iv;                                            // dlogp/dmu = (a-mu)/sigma^2
          acc_dsig[c][e] += g_logp * (d * d * iv - 1.0f) / s_sigma[j];        // dlogp/dsigma = ((a-mu)^2 - sigma^2)/sigma^3
        }
        __stcs(reinterpret_cast<float4*>(p.grad_mu + row) + ch, make_float4(g[0], g[1], g[2], g[3]));
      }
    }
  }

  // ---- block reduction in a fixed order
  acc_surr = lt::warp_sum(acc_surr);
  acc_vloss = lt::warp_sum(acc_vloss);
  acc_kl = lt::warp_sum(acc_kl);
#pragma unroll
  for (int c = 0; c < kChunks; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = acc_dsig[c][e];
      v += __shfl_xor_sync(LT_FULL_MASK, v, 4);
      v += __shfl_xor_sync(LT_FULL_MASK, v, 8);
      v += __shfl_xor_sync(LT_FULL_MASK, v, 16);
      const int j = 4 * (sub + 4 * c) + e;
      if (lane < 4 && j < A) s_red[warp][3 + j] = v;
    }
  if (lane == 0) {
    s_red[warp][0] = acc_surr;
    s_red[warp][1] = acc_vloss;
    s_red[warp][2] = acc_kl;
  }
  fold_and_finalize(p, s_red, s_sigma, s_log_sigma, inv_b);
}

__global__ void adaptive_lr_kernel(const float* kl_sum, float kl_scale, float desired_kl, float* lr_inout) {
  const float kl = *kl_sum * kl_scale;
  float lr = *lr_inout;
  if (kl > desired_kl * 2.0f)
    lr = fmaxf(1e-5f, lr / 1.5f);
  else if (kl < desired_kl / 2.0f && kl > 0.0f)
    lr = fminf(1e-2f, lr * 1.5f);
  *lr_inout = lr;
}

int grid_for(int B) {
  const int want = (int)lt::ceil_div(B, kThreads / 4);
  const int cap = 4 * lt::sm_count();
  return want < cap ? want : cap;
}

}  // namespace

extern "C" int64_t lt_ppo_loss_workspace_bytes(int B, int A) {
  const int64_t blocks = lt::ceil_div(B > 0 ? B : 1, kThreads / 4);
  return 16 + blocks * (3 + (int64_t)A) * (int64_t)sizeof(float);
}

extern "C" int lt_ppo_loss(const LtPpoLossArgs* a, void* stream) {
  if (!a || a->B <= 0 || a->A <= 0 || (a->A & 3) || a->A > kMaxA) return LT_ERR_INVALID_ARG;
  if (!a->mu || !a->sigma || !a->value || !a->actions || !a->old_logp || !a->old_mu || !a->old_sigma || !a->advantages ||
      !a->returns || !a->old_values || !a->grad_mu || !a->grad_value || !a->grad_sigma || !a->out || !a->workspace)
    return LT_ERR_INVALID_ARG;
  const uintptr_t align = (uintptr_t)a->mu | (uintptr_t)a->actions | (uintptr_t)a->old_mu | (uintptr_t)a->old_sigma | (uintptr_t)a->grad_mu;
  if (align & 15) return LT_ERR_INVALID_ARG;
  const int grid = grid_for(a->B);
  if (a->workspace_bytes < 16 + (int64_t)grid * (3 + a->A) * (int64_t)sizeof(float)) return LT_ERR_WORKSPACE;
  Params p;
  p.B = a->B; p.A = a->A;
  p.mu = a->mu; p.sigma = a->sigma; p.value = a->value; p.actions = a->actions; p.old_logp = a->old_logp;
  p.old_mu = a->old_mu; p.old_sigma = a->old_sigma; p.adv = a->advantages; p.returns = a->returns; p.old_values = a->old_values;
  p.clip = a->clip_param;
  p.clip_lo = (float)(1.0 - (double)a->clip_param);
  p.clip_hi = (float)(1.0 + (double)a->clip_param);
  p.vcoef = a->value_loss_coef; p.ecoef = a->entropy_coef; p.use_clipped_value = a->use_clipped_value_loss;
  p.desired_kl = a->desired_kl; p.grad_scale = a->grad_scale;
  p.grad_mu = a->grad_mu; p.grad_value = a->grad_value; p.grad_sigma = a->grad_sigma; p.out = a->out;
  p.lr_inout = a->lr_inout; p.loss_accum = a->loss_accum; p.ws = (PpoWs*)a->workspace;
  const int per_lane = (a->A / 4 + 3) / 4;
  cudaStream_t st = (cudaStream_t)stream;
  switch (per_lane) {
    case 1: ppo_loss_kernel<1><<<grid, kThreads, 0, st>>>(p); break;
    case 2: ppo_loss_kernel<2><<<grid, kThreads, 0, st>>>(p); break;
    case 3: ppo_loss_kernel<3><<<grid, kThreads, 0, st>>>(p); break;
    default: ppo_loss_kernel<4><<<grid, kThreads, 0, st>>>(p); break;
  }
  return lt::check_launch();
}

extern "C" int lt_adaptive_lr(const float* kl_sum, float kl_scale, float desired_kl, float* lr_inout, void* stream) {
  if (!kl_sum || !lr_inout) return LT_ERR_INVALID_ARG;
  adaptive_lr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(kl_sum, kl_scale, desired_kl, lr_inout);
  return lt::check_launch();
}
