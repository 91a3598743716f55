// Shared pieces of the PPO loss kernels (K6 `ppo_loss.cu`, K16 `ppo_heads.cu`): parameter block, workspace layout, the per-sample
// loss / gradient expressions of reference loco_rl/loco_rl/algorithms/ppo.py:252-302 and the deterministic batch reduction
// (block partial rows -> the last block folds them in index order, writes the means, dL/dsigma and the adaptive learning rate).
#pragma once
#include "lt_common.cuh"

namespace lt_ppo {

constexpr int kThreads = 256;
constexpr int kMaxChunks = 4;   // A <= 64
constexpr int kMaxA = 64;
constexpr float kHalfLog2Pi = 0.91893853320467274178f;  // log(sqrt(2 pi))

struct PpoWs {
  unsigned int counter;
  unsigned int pad[3];
  float partial[1];  // [blocks][3 + A]
};

struct Params {
  int B, A;
  const float *mu, *sigma, *value, *actions, *old_logp, *old_mu, *old_sigma, *adv, *returns, *old_values;
  float clip, clip_lo, clip_hi, vcoef, ecoef;
  int use_clipped_value;
  float desired_kl, grad_scale;
  float *grad_mu, *grad_value, *grad_sigma, *out, *lr_inout, *loss_accum;
  PpoWs* ws;
};

// surrogate (ppo.py:284-289) with its derivative w.r.t. the log-prob, value loss (ppo.py:292-300) with its derivative w.r.t. the value
struct SampleTerms {
  float surr, g_logp, vloss, g_val;
};
__device__ __forceinline__ SampleTerms sample_terms(const Params& p, float logp, float old_logp, float adv, float val, float old_val, float ret, float inv_b) {
  SampleTerms t;
  const float ratio = expf(logp - old_logp);
  const float s_un = -adv * ratio;
  const float s_cl = -adv * fminf(fmaxf(ratio, p.clip_lo), p.clip_hi);
  t.surr = fmaxf(s_un, s_cl);
  t.g_logp = (s_un >= s_cl ? s_un : 0.f) * inv_b * p.grad_scale;  // d/dlogp(-A*ratio) = -A*ratio
  const float e1 = val - ret;
  if (p.use_clipped_value) {
    const float d = val - old_val;
    const float dc = fminf(fmaxf(d, -p.clip), p.clip);
    const float e2 = (old_val + dc) - ret;
    const float v1 = e1 * e1, v2 = e2 * e2;
    t.vloss = fmaxf(v1, v2);
    const float in_rng = (d >= -p.clip && d <= p.clip) ? 1.f : 0.f;
    t.g_val = v1 > v2 ? 2.f * e1 : (v1 < v2 ? 2.f * e2 * in_rng : e1 + e2 * in_rng);
  } else {
    t.vloss = e1 * e1;
    t.g_val = 2.f * e1;
  }
  t.g_val *= p.vcoef * inv_b * p.grad_scale;
  return t;
}

// Block partial row -> workspace; the last block to arrive folds all rows in index order and finalises.  `s_red[w][k]` holds warp w's
// partial of column k (0 surrogate, 1 value loss, 2 KL, 3 + j dL/dsigma_j); s_sigma / s_log_sigma are the block's copies of sigma.
__device__ __forceinline__ void fold_and_finalize(const Params& p, float (*s_red)[3 + kMaxA], const float* s_sigma, const float* s_log_sigma, float inv_b) {
  __shared__ bool is_last;
  const int A = p.A, K = 3 + A;
  __syncthreads();
  float* my_partial = p.ws->partial + (size_t)blockIdx.x * K;
  for (int k = threadIdx.x; k < K; k += kThreads) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) v += s_red[w][k];
    my_partial[k] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&p.ws->counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;

  // ---- finalize (one block): fold block partials in index order
  __threadfence();
  // Work item = (row group g of kFoldGroups, column k): a thread adds rows g, g + kFoldGroups, ... of its column -- consecutive
  // threads read consecutive floats of a partial row, and the loads of an unrolled batch are in flight together.  (One warp
  // per column with 12 dependent L2 round trips per lane, two columns per warp, was half of this kernel's 13.8 us.)
  constexpr int kFoldGroups = 16;
  __shared__ float s_tot[3 + kMaxA];
  float* s_fold = &s_red[0][0];  // reused: kFoldGroups x K <= (kThreads / 32) x (3 + kMaxA) needs kFoldGroups <= ... see static_assert
  static_assert(kFoldGroups * 15 <= (kThreads / 32) * (3 + kMaxA), "fold scratch (A = 12) must fit the reduction scratch");
  const int groups = (kFoldGroups * K <= (kThreads / 32) * (3 + kMaxA)) ? kFoldGroups : (kThreads / 32) * (3 + kMaxA) / K;
  for (int idx = threadIdx.x; idx < groups * K; idx += kThreads) {
    const int g = idx / K, k = idx - g * K;
    // eight loads per trip with no bounds test between them (a test per load makes every load wait for the previous add);
    // fixed association order, so the result does not depend on timing
    const float* col = p.ws->partial + k;
    const int G = (int)gridDim.x;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int i = g;
    for (; i + 7 * groups < G; i += 8 * groups) {
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = __ldcg(col + (size_t)(i + u * groups) * K);
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += x[u];
    }
    for (; i < G; i += groups) a[0] += __ldcg(col + (size_t)i * K);
    s_fold[idx] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += kThreads) {
    float v = 0.f;
    for (int g = 0; g < groups; ++g) v += s_fold[g * K + k];
    s_tot[k] = v;
  }
  __syncthreads();
  if (threadIdx.x < A) {
    const int j = threadIdx.x;
    // d(-ecoef * mean(entropy))/dsigma_j = -ecoef / sigma_j
    p.grad_sigma[j] = s_tot[3 + j] - p.ecoef * p.grad_scale / s_sigma[j];
  }
  if (threadIdx.x == 0) {
    float entropy = 0.f;
    for (int j = 0; j < A; ++j) entropy += 0.5f + kHalfLog2Pi + s_log_sigma[j];  // Normal.entropy().sum(-1)
    const float surr = s_tot[0] * inv_b, vloss = s_tot[1] * inv_b, kl = s_tot[2] * inv_b;
    p.out[0] = surr + p.vcoef * vloss - p.ecoef * entropy;  // ppo.py:302
    p.out[1] = surr;
    p.out[2] = vloss;
    p.out[3] = entropy;
    p.out[4] = kl;
    float lr = p.lr_inout ? *p.lr_inout : 0.f;
    if (p.lr_inout && p.desired_kl > 0.f) {  // ppo.py:275-281
      if (kl > p.desired_kl * 2.0f)
        lr = fmaxf(1e-5f, lr / 1.5f);
      else if (kl < p.desired_kl / 2.0f && kl > 0.0f)
        lr = fminf(1e-2f, lr * 1.5f);
      *p.lr_inout = lr;
    }
    p.out[5] = lr;
    p.out[6] = 0.f;
    p.out[7] = 0.f;
    if (p.loss_accum) {  // ppo.py:361-363 without the three .item() syncs
      p.loss_accum[0] += vloss;
      p.loss_accum[1] += surr;
      p.loss_accum[2] += entropy;
      p.loss_accum[3] += 1.0f;
    }
    p.ws->counter = 0;  // self-cleaning
  }
}

inline int loss_grid_cap() { return 4 * lt::sm_count(); }

}  // namespace lt_ppo
