// K16: the output heads of both MLPs, the PPO loss and the head dgrad of the backward pass, in ONE pass.
// Per mini-batch it replaces, of reference loco_rl/loco_rl/algorithms/ppo.py:252-302,350 under one process:
//   * the two head layers  mu = h_a W_a^T + b_a  (A x H),  V = h_c W_c^T + b_c  (1 x H)  (modules/actor_critic.py:33-56; cuBLAS
//     GEMM / GEMV launches in round 1),
//   * the loss block itself (K6: log-prob, KL, surrogate, value loss, their analytic derivatives, the adaptive learning rate),
//   * autograd's head dgrad  g_h = (dL/dout . W) * elu'(h)  for both networks (two cuBLAS GEMMs + two K9 passes),
// i.e. seven launches that each stream the [B, H] hidden activations or gradients once more.  Here a row of h_a and h_c is read once
// and its two gradient rows are written once: (2 x H x 4 B read + 2 x H x 4 B written + 264 B of rollout row + (A + 1) x 4 B of
// dL/dmu, dL/dV for the head weight gradients, which K15 takes on the trailing streams) per sample.
//
// Mapping: one warp per sample, lane l owns columns [4 l v, 4 l v + 4 v) of the hidden row (v = H / 128 float4 per lane); the head
// weights of those columns live in registers for the whole launch.  The A <= 16 dot products of a row are reduced by a TRANSPOSED
// butterfly (16 shuffles instead of 16 x 5): after it lane l holds mu_{l >> 1}; that lane evaluates the per-action terms of the loss,
// the per-sample sums take two more butterflies, and the A gradient values return to all lanes by A shuffles for the dgrad.  The
// next row of a warp is in flight while the current one is processed.  Loss statistics take the deterministic two-stage path of K6.
#include "ppo_loss_common.cuh"

namespace {

using namespace lt_ppo;

constexpr int kHeadA = 16;      // action slots of the transposed reduction (A <= 16, A % 4 == 0)
constexpr int kWarps = kThreads / 32;

struct HeadParams {
  Params loss;                  // loss.mu / loss.value / loss.grad_mu / loss.grad_value: optional OUTPUTS here (may be null)
  int H;                        // width of the last hidden layer of both networks (multiple of 128)
  const float *h_actor, *h_critic, *w_actor, *b_actor, *w_critic, *b_critic;
  float *g_h_actor, *g_h_critic;      // [B, H] dLoss/d(pre-activation of the last hidden layer)
};

__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}

// 16 per-lane partials -> lane l holds the warp total of slot l >> 1 (both lanes of a pair hold it)
__device__ __forceinline__ float transposed_reduce16(const float (&v)[kHeadA], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  float w[8], x[4], y[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? v[i] : v[i + 8], keep = b4 ? v[i + 8] : v[i];
    w[i] = keep + __shfl_xor_sync(LT_FULL_MASK, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? w[i] : w[i + 4], keep = b3 ? w[i + 4] : w[i];
    x[i] = keep + __shfl_xor_sync(LT_FULL_MASK, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? x[i] : x[i + 2], keep = b2 ? x[i + 2] : x[i];
    y[i] = keep + __shfl_xor_sync(LT_FULL_MASK, send, 4);
  }
  const float send = b1 ? y[0] : y[1], keep = b1 ? y[1] : y[0];
  float z = keep + __shfl_xor_sync(LT_FULL_MASK, send, 2);
  z += __shfl_xor_sync(LT_FULL_MASK, z, 1);
  return z;
}

template <int V, int NA>  // V float4 per lane of a hidden row (H = 128 V); NA = action slots kept in registers (A <= NA)
__global__ void __launch_bounds__(kThreads, V == 1 ? 2 : 1) ppo_heads_kernel(const HeadParams hp) {
  const Params& p = hp.loss;
  __shared__ float s_sigma[kMaxA], s_inv_var[kMaxA], s_log_sigma[kMaxA];
  __shared__ float s_red[kWarps][3 + kMaxA];
  const int A = p.A, H = hp.H;
  for (int j = threadIdx.x; j < A; j += kThreads) {
    const float s = p.sigma[j];
    s_sigma[j] = s;
    s_inv_var[j] = 1.0f / (s * s);
    s_log_sigma[j] = logf(s);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j_own = lane >> 1;                        // the action this lane evaluates after the transposed reduction
  const bool own = !(lane & 1) && j_own < A;          // even lane of the pair, live action
  const float inv_b = 1.0f / (float)p.B;
  const float sig = own ? s_sigma[j_own] : 1.0f, iv = own ? s_inv_var[j_own] : 0.0f, lsig = own ? s_log_sigma[j_own] : 0.0f;

  // head weights of this lane's columns
  float4 wa[NA][V], wc[V];
#pragma unroll
  for (int j = 0; j < NA; ++j)
#pragma unroll
    for (int v = 0; v < V; ++v)
      wa[j][v] = j < A ? __ldg(reinterpret_cast<const float4*>(hp.w_actor + (size_t)j * H) + lane * V + v) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int v = 0; v < V; ++v) wc[v] = __ldg(reinterpret_cast<const float4*>(hp.w_critic) + lane * V + v);
  const float bias_a = own ? __ldg(hp.b_actor + j_own) : 0.0f;
  const float bias_c = __ldg(hp.b_critic);

  float acc_surr = 0.f, acc_vloss = 0.f, acc_kl = 0.f, acc_dsig = 0.f;

  const int warps_total = gridDim.x * kWarps;
  int b = blockIdx.x * kWarps + warp;                  // warp-uniform
  // the warp's NEXT row (hidden rows and the rollout scalars) travels while the current one is processed
  struct RowScalars {
    float act, omu, osg, old_logp, adv, old_val, ret;
  };
  auto load_scalars = [&](int r) {
    RowScalars q;
    q.act = 0.f; q.omu = 0.f; q.osg = 1.f;
    if (own) {
      const size_t rr = (size_t)r * A + j_own;
      q.act = __ldcs(p.actions + rr);
      q.omu = __ldcs(p.old_mu + rr);
      q.osg = __ldcs(p.old_sigma + rr);
    }
    q.old_logp = __ldcs(p.old_logp + r);
    q.adv = __ldcs(p.adv + r);
    q.old_val = __ldcs(p.old_values + r);
    q.ret = __ldcs(p.returns + r);
    return q;
  };
  float4 ha[V], hc[V], ha_n[V], hc_n[V];
  RowScalars sc_n;
  sc_n.act = sc_n.omu = sc_n.old_logp = sc_n.adv = sc_n.old_val = sc_n.ret = 0.f;
  sc_n.osg = 1.f;
  if (b < p.B) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      ha_n[v] = __ldcs(reinterpret_cast<const float4*>(hp.h_actor + (size_t)b * H) + lane * V + v);
      hc_n[v] = __ldcs(reinterpret_cast<const float4*>(hp.h_critic + (size_t)b * H) + lane * V + v);
    }
    sc_n = load_scalars(b);
  }
  for (; b < p.B; b += warps_total) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      ha[v] = ha_n[v];
      hc[v] = hc_n[v];
    }
    const RowScalars sc = sc_n;
    const int bn = b + warps_total;
    if (bn < p.B) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        ha_n[v] = __ldcs(reinterpret_cast<const float4*>(hp.h_actor + (size_t)bn * H) + lane * V + v);
        hc_n[v] = __ldcs(reinterpret_cast<const float4*>(hp.h_critic + (size_t)bn * H) + lane * V + v);
      }
      sc_n = load_scalars(bn);
    }
    const size_t row = (size_t)b * A;
    const float act = sc.act, omu = sc.omu, osg = sc.osg, old_logp = sc.old_logp, adv = sc.adv, old_val = sc.old_val, ret = sc.ret;

    // ---- heads
    float part[kHeadA];
#pragma unroll
    for (int j = 0; j < kHeadA; ++j) {
      float s = 0.f;
      if (j < NA) {
#pragma unroll
        for (int v = 0; v < V; ++v) s = dot4(ha[v], wa[j][v], s);
      }
      part[j] = s;
    }
    const float mu = transposed_reduce16(part, lane) + bias_a;    // mu_{j_own} (0 for the dead slots)
    float vs = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) vs = dot4(hc[v], wc[v], vs);
    const float val = lt::warp_sum(vs) + bias_c;

    // ---- loss (ppo.py:252-302): per-action terms in the owning lanes, per-sample sums by butterflies
    const float d = act - mu;
    float logp = 0.f, kl = 0.f;
    if (own) {
      const float half_iv = 0.5f * iv;
      logp = -(d * d) * half_iv - lsig - kHalfLog2Pi;
      const float dm = omu - mu;
      kl = logf(sig / osg + 1.0e-5f) + (osg * osg + dm * dm) * half_iv - 0.5f;
    }
    logp = lt::warp_sum(logp);
    kl = lt::warp_sum(kl);
    const SampleTerms t = sample_terms(p, logp, old_logp, adv, val, old_val, ret, inv_b);
    const float g_mu = own ? t.g_logp * d * iv : 0.f;             // dlogp/dmu = (a-mu)/sigma^2
    if (own) {
      acc_dsig += t.g_logp * (d * d * iv - 1.0f) / sig;            // dlogp/dsigma = ((a-mu)^2 - sigma^2)/sigma^3
      p.grad_mu[row + j_own] = g_mu;
      if (p.mu) const_cast<float*>(p.mu)[row + j_own] = mu;
    }
    if (lane == 0) {
      acc_surr += t.surr;
      acc_vloss += t.vloss;
      acc_kl += kl;
      p.grad_value[b] = t.g_val;
      if (p.value) const_cast<float*>(p.value)[b] = val;
    }

    // ---- backward through the heads: dgrad with the ELU backward of the last hidden layer in the same registers
    float4 ga[V];
#pragma unroll
    for (int v = 0; v < V; ++v) ga[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      const float gj = __shfl_sync(LT_FULL_MASK, g_mu, 2 * j);    // 0 for j >= A
#pragma unroll
      for (int v = 0; v < V; ++v) {
        ga[v].x = fmaf(gj, wa[j][v].x, ga[v].x); ga[v].y = fmaf(gj, wa[j][v].y, ga[v].y);
        ga[v].z = fmaf(gj, wa[j][v].z, ga[v].z); ga[v].w = fmaf(gj, wa[j][v].w, ga[v].w);
      }
    }
    const float gv = t.g_val;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      // elu'(x) from the stored post-activation h (alpha = 1): h > 0 ? 1 : h + 1
      float4 o;
      o.x = ga[v].x * (ha[v].x > 0.f ? 1.f : ha[v].x + 1.f); o.y = ga[v].y * (ha[v].y > 0.f ? 1.f : ha[v].y + 1.f);
      o.z = ga[v].z * (ha[v].z > 0.f ? 1.f : ha[v].z + 1.f); o.w = ga[v].w * (ha[v].w > 0.f ? 1.f : ha[v].w + 1.f);
      __stcs(reinterpret_cast<float4*>(hp.g_h_actor + (size_t)b * H) + lane * V + v, o);
      float4 c;
      c.x = gv * wc[v].x * (hc[v].x > 0.f ? 1.f : hc[v].x + 1.f); c.y = gv * wc[v].y * (hc[v].y > 0.f ? 1.f : hc[v].y + 1.f);
      c.z = gv * wc[v].z * (hc[v].z > 0.f ? 1.f : hc[v].z + 1.f); c.w = gv * wc[v].w * (hc[v].w > 0.f ? 1.f : hc[v].w + 1.f);
      __stcs(reinterpret_cast<float4*>(hp.g_h_critic + (size_t)b * H) + lane * V + v, c);
    }
  }

  // ---- loss statistics: per-warp partial rows in the layout K6 uses, then the shared deterministic fold
  if (lane == 0) {
    s_red[warp][0] = acc_surr;
    s_red[warp][1] = acc_vloss;
    s_red[warp][2] = acc_kl;
  }
  if (own) s_red[warp][3 + j_own] = acc_dsig;
  fold_and_finalize(p, s_red, s_sigma, s_log_sigma, inv_b);
}

// ------------------------------------------------------------------------------------------------------------------------------
// K3b: the same heads on the ROLLOUT side -- mu = h_a W_a^T + b_a, V = h_c W_c^T + b_c, a = mu + sigma * eps, log-prob -- one launch per
// env step in place of a cuBLAS GEMM, a GEMV, the sample kernel (K3) and the value copy (reference modules/actor_critic.py:105-131 as
// driven by algorithms/ppo.py:129-141).  Same mapping as above (warp per env, transposed butterfly); the normal draws are the Philox
// stream of lt_act_sample (key (seed, offset), counter (env, chunk)), or explicit eps.
struct ActHeadParams {
  int N, A, H;
  const float *h_actor, *h_critic, *w_actor, *b_actor, *w_critic, *b_critic, *sigma, *eps;
  float *actions, *logp, *mu_out, *sigma_out, *values;
  uint64_t seed, offset;
  const int64_t* offset_base;
};

template <int V, int NA>
__global__ void __launch_bounds__(kThreads) act_heads_kernel(const ActHeadParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int A = p.A, H = p.H;
  const int j_own = lane >> 1;
  const bool own = !(lane & 1) && j_own < A;
  uint64_t offset = p.offset;
  if (p.offset_base) offset += (uint64_t)*p.offset_base;
  float4 wa[NA][V], wc[V];
#pragma unroll
  for (int j = 0; j < NA; ++j)
#pragma unroll
    for (int v = 0; v < V; ++v)
      wa[j][v] = j < A ? __ldg(reinterpret_cast<const float4*>(p.w_actor + (size_t)j * H) + lane * V + v) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int v = 0; v < V; ++v) wc[v] = p.h_critic ? __ldg(reinterpret_cast<const float4*>(p.w_critic) + lane * V + v) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float bias_a = own ? __ldg(p.b_actor + j_own) : 0.f;
  const float bias_c = p.h_critic ? __ldg(p.b_critic) : 0.f;
  const float sig = own ? __ldg(p.sigma + j_own) : 1.f;
  const float lsig = logf(sig);
  for (int n = blockIdx.x * kWarps + warp; n < p.N; n += gridDim.x * kWarps) {
    float part[kHeadA];
    float4 ha[V];
#pragma unroll
    for (int v = 0; v < V; ++v) ha[v] = __ldcs(reinterpret_cast<const float4*>(p.h_actor + (size_t)n * H) + lane * V + v);
    float vs = 0.f;
    if (p.h_critic) {
#pragma unroll
      for (int v = 0; v < V; ++v) vs = dot4(__ldcs(reinterpret_cast<const float4*>(p.h_critic + (size_t)n * H) + lane * V + v), wc[v], vs);
    }
#pragma unroll
    for (int j = 0; j < kHeadA; ++j) {
      float s = 0.f;
      if (j < NA) {
#pragma unroll
        for (int v = 0; v < V; ++v) s = dot4(ha[v], wa[j][v], s);
      }
      part[j] = s;
    }
    const float mu = transposed_reduce16(part, lane) + bias_a;
    const float val = lt::warp_sum(vs) + bias_c;
    float lp = 0.f;
    if (own) {
      const size_t o = (size_t)n * A + j_own;
      float e;
      if (p.eps) {
        e = __ldcs(p.eps + o);
      } else {
        const uint4 r = lt::Philox::gen(p.seed, offset, (uint32_t)n, (uint32_t)(j_own >> 2));
        const int k = j_own & 3;
        const float2 g = k < 2 ? lt::Philox::normal2(r.x, r.y) : lt::Philox::normal2(r.z, r.w);
        e = (k & 1) ? g.y : g.x;
      }
      const float a = __fadd_rn(mu, __fmul_rn(sig, e));   // torch.normal(mean, std): mean + std * eps
      const float d = a - mu;
      lp = -(d * d) / (2.0f * sig * sig) - lsig - kHalfLog2Pi;
      p.actions[o] = a;
      if (p.mu_out) p.mu_out[o] = mu;
      if (p.sigma_out) p.sigma_out[o] = sig;
    }
    lp = lt::warp_sum(lp);
    if (lane == 0) {
      p.logp[n] = lp;
      if (p.values) p.values[n] = val;
    }
  }
}

}  // namespace

extern "C" int64_t lt_ppo_heads_workspace_bytes(int B, int A) {
  (void)B;
  return 16 + (int64_t)lt_ppo::loss_grid_cap() * (3 + (int64_t)A) * (int64_t)sizeof(float);
}

extern "C" int lt_ppo_heads_loss(const LtPpoHeadsArgs* h, void* stream) {
  if (!h) return LT_ERR_INVALID_ARG;
  const LtPpoLossArgs* a = &h->loss;
  if (a->B <= 0 || a->A <= 0) return LT_ERR_INVALID_ARG;
  if ((a->A & 3) || a->A > kHeadA || h->H <= 0 || (h->H % 128) != 0 || h->H > 512) return LT_ERR_UNSUPPORTED;
  if (!a->sigma || !a->actions || !a->old_logp || !a->old_mu || !a->old_sigma || !a->advantages || !a->returns || !a->old_values ||
      !a->grad_mu || !a->grad_value || !a->grad_sigma || !a->out || !a->workspace)
    return LT_ERR_INVALID_ARG;
  if (!h->h_actor || !h->h_critic || !h->w_actor || !h->b_actor || !h->w_critic || !h->b_critic || !h->g_h_actor || !h->g_h_critic)
    return LT_ERR_INVALID_ARG;
  const uintptr_t align = (uintptr_t)h->h_actor | (uintptr_t)h->h_critic | (uintptr_t)h->w_actor | (uintptr_t)h->w_critic | (uintptr_t)h->g_h_actor |
                          (uintptr_t)h->g_h_critic;
  if (align & 15) return LT_ERR_INVALID_ARG;
  const int want = (int)lt::ceil_div(a->B, kWarps);
  const int cap = 2 * lt::sm_count();   // two co-resident blocks per SM, every warp loops over its rows
  const int grid = want < cap ? want : cap;
  if (a->workspace_bytes < 16 + (int64_t)grid * (3 + a->A) * (int64_t)sizeof(float)) return LT_ERR_WORKSPACE;
  HeadParams hp;
  Params& p = hp.loss;
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
  hp.H = h->H;
  hp.h_actor = h->h_actor; hp.h_critic = h->h_critic; hp.w_actor = h->w_actor; hp.b_actor = h->b_actor;
  hp.w_critic = h->w_critic; hp.b_critic = h->b_critic; hp.g_h_actor = h->g_h_actor; hp.g_h_critic = h->g_h_critic;
  cudaStream_t st = (cudaStream_t)stream;
  const int v = h->H / 128, na = a->A <= 4 ? 4 : (a->A <= 8 ? 8 : (a->A <= 12 ? 12 : 16));
#define LT_HEADS_CASE(V_, NA_) \
  if (v == V_ && na == NA_) { ppo_heads_kernel<V_, NA_><<<grid, kThreads, 0, st>>>(hp); return lt::check_launch(); }
  LT_HEADS_CASE(1, 4) LT_HEADS_CASE(1, 8) LT_HEADS_CASE(1, 12) LT_HEADS_CASE(1, 16)
  LT_HEADS_CASE(2, 4) LT_HEADS_CASE(2, 8) LT_HEADS_CASE(2, 12) LT_HEADS_CASE(2, 16)
#undef LT_HEADS_CASE
  return LT_ERR_UNSUPPORTED;
}

extern "C" int lt_act_heads(const float* h_actor, const float* h_critic, const float* w_actor, const float* b_actor, const float* w_critic,
                            const float* b_critic, const float* sigma, const float* eps, float* actions, float* logp, float* mu_out,
                            float* sigma_out, float* values, int N, int A, int H, uint64_t seed, uint64_t offset, const int64_t* offset_base,
                            void* stream) {
  if (!h_actor || !w_actor || !b_actor || !sigma || !actions || !logp || N <= 0 || A <= 0) return LT_ERR_INVALID_ARG;
  if (h_critic && (!w_critic || !b_critic || !values)) return LT_ERR_INVALID_ARG;
  if ((A & 3) || A > kHeadA || H <= 0 || (H % 128) != 0 || H > 256) return LT_ERR_UNSUPPORTED;
  if ((((uintptr_t)h_actor | (uintptr_t)h_critic | (uintptr_t)w_actor | (uintptr_t)w_critic) & 15) != 0) return LT_ERR_INVALID_ARG;
  ActHeadParams p;
  p.N = N; p.A = A; p.H = H;
  p.h_actor = h_actor; p.h_critic = h_critic; p.w_actor = w_actor; p.b_actor = b_actor; p.w_critic = w_critic; p.b_critic = b_critic;
  p.sigma = sigma; p.eps = eps; p.actions = actions; p.logp = logp; p.mu_out = mu_out; p.sigma_out = sigma_out; p.values = values;
  p.seed = seed; p.offset = offset; p.offset_base = offset_base;
  const int want = (int)lt::ceil_div(N, kWarps);
  const int cap = 4 * lt::sm_count();
  const int grid = want < cap ? want : cap;
  cudaStream_t st = (cudaStream_t)stream;
  const int v = H / 128, na = A <= 4 ? 4 : (A <= 8 ? 8 : (A <= 12 ? 12 : 16));
#define LT_ACT_CASE(V_, NA_) \
  if (v == V_ && na == NA_) { act_heads_kernel<V_, NA_><<<grid, kThreads, 0, st>>>(p); return lt::check_launch(); }
  LT_ACT_CASE(1, 4) LT_ACT_CASE(1, 8) LT_ACT_CASE(1, 12) LT_ACT_CASE(1, 16)
  LT_ACT_CASE(2, 4) LT_ACT_CASE(2, 8) LT_ACT_CASE(2, 12) LT_ACT_CASE(2, 16)
#undef LT_ACT_CASE
  return LT_ERR_UNSUPPORTED;
}
