// K17: the tactile pre-encoder of the CNN-RNN student -- conv1 + ReLU + maxpool + conv2 + ReLU + conv3 + ReLU + flatten + head
// Linear -- as ONE kernel per batch of tactile frames (inference path: the student acting in the DAgger collection / evaluation).
// Replaces reference loco_rl/loco_rl/models/cnn_2d.py:16-131 (`CNN2dHead` = `CNN2d` + `MLP` head) as instantiated by
// locotouch/config/locotouch/agents/distillation_cfg.py:78-85 with loco_rl/models/model_cfg.py:17-25: image (2, 17, 13), channels
// (24, 24, 24), kernels (4, 3, 2), use_maxpool with strides (2, 1, 1) -> MaxPool2d(2) after conv1, ReLU, no padding, head Linear
// 192 -> E -- called from locotouch/distill/student.py:88-103.  The reference runs 3 cuDNN convolutions, 3 ReLU, a max-pool, a
// reshape and a cuBLAS GEMM per call (8 launches whose intermediates go through HBM); here a frame's activations never leave the SM.
//
// Input: either the fp32 observation [M, 2*17*13] or the ballot-packed bitmap [M, 7] that K2 emits (bit t%32 of word t/32 = taxel t,
// both channels identical: the student never needs the 442-float image).
// Mapping: one warp per frame, lane = output channel (24 of 32 lanes carry channels); a lane keeps its channel's outputs of a layer
// in registers and the warp exchanges layers through a private shared-memory scratch (1.9 K floats).  conv2 / conv3 / head weights
// sit in shared memory transposed to [in][tap][out] so that the lanes of a warp read consecutive words; conv1 weights (32 per
// channel) live in registers.  Per input plane a lane loads the plane once (broadcast reads) and applies all taps to all of its
// output positions: 135 FMA per 44 shared loads in conv2.
#include "lt_common.cuh"

namespace {

constexpr int C0 = 2, H0 = 17, W0 = 13, IMG = C0 * H0 * W0;   // 442
constexpr int C1 = 24, K1 = 4, H1 = H0 - K1 + 1, W1 = W0 - K1 + 1;  // 14 x 10
constexpr int HP = H1 / 2, WP = W1 / 2;                        // 7 x 5 after MaxPool2d(2)
constexpr int C2 = 24, K2 = 3, H2 = HP - K2 + 1, W2 = WP - K2 + 1;  // 5 x 3
constexpr int C3 = 24, K3 = 2, H3 = H2 - K3 + 1, W3 = W2 - K3 + 1;  // 4 x 2
constexpr int F = C3 * H3 * W3;                                // 192
constexpr int EMAX = 64;
constexpr int kWarps = 12, kThreads = kWarps * 32;
constexpr int kScratch = 448 + C1 * HP * WP + C2 * H2 * W2 + F;  // img (padded) + pooled + conv2 out + conv3 out = 1840 floats

struct CnnParams {
  const float* image;       // [M, 442] or null
  const uint32_t* packed;   // [M, words] or null
  int words, M, E;
  const float *w1, *b1, *w2, *b2, *w3, *b3, *wh, *bh;
  float* out;               // [M, E]
};

__global__ void __launch_bounds__(kThreads, 1) student_cnn_kernel(const CnnParams p) {
  extern __shared__ float smem[];
  float* sW2 = smem;                              // [C1][9][C2]
  float* sW3 = sW2 + C1 * K2 * K2 * C2;           // [C2][4][C3]
  float* sWh = sW3 + C2 * K3 * K3 * C3;           // [F][E]
  float* sB = sWh + F * EMAX;                     // b2[24] b3[24] bh[64]
  float* scratch = sB + 128;
  const int E = p.E;
  for (int i = threadIdx.x; i < C2 * C1 * K2 * K2; i += kThreads) {   // w2 [oc][ic][tap] -> [ic][tap][oc]
    const int oc = i / (C1 * 9), r = i - oc * (C1 * 9);
    sW2[r * C2 + oc] = p.w2[i];
  }
  for (int i = threadIdx.x; i < C3 * C2 * K3 * K3; i += kThreads) {   // w3 [oc][ic][tap] -> [ic][tap][oc]
    const int oc = i / (C2 * 4), r = i - oc * (C2 * 4);
    sW3[r * C3 + oc] = p.w3[i];
  }
  for (int i = threadIdx.x; i < E * F; i += kThreads) {               // wh [o][i] -> [i][o]
    const int o = i / F, k = i - o * F;
    sWh[k * EMAX + o] = p.wh[i];
  }
  for (int i = threadIdx.x; i < 128; i += kThreads) sB[i] = i < 24 ? p.b2[i] : (i < 48 ? p.b3[i - 24] : (i - 64 >= 0 && i - 64 < E ? p.bh[i - 64] : 0.f));
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool chan = lane < C1;                    // C1 == C2 == C3 == 24 channel lanes
  float* img = scratch + warp * kScratch;         // [2][17][13]
  float* pooled = img + 448;                      // [24][35]
  float* c2 = pooled + C1 * HP * WP;              // [24][15]
  float* c3 = c2 + C2 * H2 * W2;                  // [192]
  float w1[C0 * K1 * K1];
#pragma unroll
  for (int k = 0; k < C0 * K1 * K1; ++k) w1[k] = chan ? __ldg(p.w1 + lane * (C0 * K1 * K1) + k) : 0.f;
  const float b1 = chan ? __ldg(p.b1 + lane) : 0.f;

  for (int m = blockIdx.x * kWarps + warp; m < p.M; m += gridDim.x * kWarps) {
    // ---- stage the frame
    if (p.packed != nullptr) {
      const uint32_t* wds = p.packed + (size_t)m * p.words;
      for (int t = lane; t < H0 * W0; t += 32) {
        const float v = (float)((__ldg(wds + (t >> 5)) >> (t & 31)) & 1u);
        img[t] = v;
        img[H0 * W0 + t] = v;
      }
    } else {
      const float* src = p.image + (size_t)m * IMG;
      for (int t = lane; t < IMG; t += 32) img[t] = __ldcs(src + t);
    }
    __syncwarp();
    // ---- conv1 (4x4, 2 -> 24) + ReLU + MaxPool2d(2): relu(max(.)) == max(0, .)
#pragma unroll 1
    for (int pp = 0; pp < HP * WP; ++pp) {
      const int py = pp / WP, px = pp - py * WP;
      float patch[C0][K1 + 1][K1 + 1];
#pragma unroll
      for (int c = 0; c < C0; ++c)
#pragma unroll
        for (int y = 0; y < K1 + 1; ++y)
#pragma unroll
          for (int x = 0; x < K1 + 1; ++x) patch[c][y][x] = img[c * (H0 * W0) + (2 * py + y) * W0 + 2 * px + x];
      float best = 0.f;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float acc = b1;
#pragma unroll
          for (int c = 0; c < C0; ++c)
#pragma unroll
            for (int ky = 0; ky < K1; ++ky)
#pragma unroll
              for (int kx = 0; kx < K1; ++kx) acc = fmaf(w1[(c * K1 + ky) * K1 + kx], patch[c][dy + ky][dx + kx], acc);
          best = fmaxf(best, acc);
        }
      if (chan) pooled[lane * (HP * WP) + pp] = best;
    }
    __syncwarp();
    // ---- conv2 (3x3, 24 -> 24) + ReLU
    {
      float acc[H2 * W2];
      const float b = chan ? sB[lane] : 0.f;
#pragma unroll
      for (int i = 0; i < H2 * W2; ++i) acc[i] = b;
#pragma unroll 1
      for (int ic = 0; ic < C1; ++ic) {
        float plane[HP * WP];
#pragma unroll
        for (int i = 0; i < HP * WP; ++i) plane[i] = pooled[ic * (HP * WP) + i];
#pragma unroll
        for (int ky = 0; ky < K2; ++ky)
#pragma unroll
          for (int kx = 0; kx < K2; ++kx) {
            const float w = chan ? sW2[(ic * 9 + ky * K2 + kx) * C2 + lane] : 0.f;
#pragma unroll
            for (int y = 0; y < H2; ++y)
#pragma unroll
              for (int x = 0; x < W2; ++x) acc[y * W2 + x] = fmaf(w, plane[(y + ky) * WP + x + kx], acc[y * W2 + x]);
          }
      }
      if (chan) {
#pragma unroll
        for (int i = 0; i < H2 * W2; ++i) c2[lane * (H2 * W2) + i] = fmaxf(acc[i], 0.f);
      }
    }
    __syncwarp();
    // ---- conv3 (2x2, 24 -> 24) + ReLU, flattened [C, H, W]
    {
      float acc[H3 * W3];
      const float b = chan ? sB[24 + lane] : 0.f;
#pragma unroll
      for (int i = 0; i < H3 * W3; ++i) acc[i] = b;
#pragma unroll 1
      for (int ic = 0; ic < C2; ++ic) {
        float plane[H2 * W2];
#pragma unroll
        for (int i = 0; i < H2 * W2; ++i) plane[i] = c2[ic * (H2 * W2) + i];
#pragma unroll
        for (int ky = 0; ky < K3; ++ky)
#pragma unroll
          for (int kx = 0; kx < K3; ++kx) {
            const float w = chan ? sW3[(ic * 4 + ky * K3 + kx) * C3 + lane] : 0.f;
#pragma unroll
            for (int y = 0; y < H3; ++y)
#pragma unroll
              for (int x = 0; x < W3; ++x) acc[y * W3 + x] = fmaf(w, plane[(y + ky) * W2 + x + kx], acc[y * W3 + x]);
          }
      }
      if (chan) {
#pragma unroll
        for (int i = 0; i < H3 * W3; ++i) c3[lane * (H3 * W3) + i] = fmaxf(acc[i], 0.f);
      }
    }
    __syncwarp();
    // ---- head Linear 192 -> E (lane owns outputs lane and lane + 32)
    {
      float a0 = sB[64 + lane], a1 = sB[96 + lane];
#pragma unroll 8
      for (int k = 0; k < F; ++k) {
        const float x = c3[k];
        a0 = fmaf(sWh[k * EMAX + lane], x, a0);
        a1 = fmaf(sWh[k * EMAX + 32 + lane], x, a1);
      }
      float* o = p.out + (size_t)m * E;
      if (lane < E) o[lane] = a0;
      if (lane + 32 < E) o[lane + 32] = a1;
    }
    __syncwarp();
  }
}

constexpr int kSmemBytes = (C1 * 9 * C2 + C2 * 4 * C3 + F * EMAX + 128 + kWarps * kScratch) * (int)sizeof(float);

}  // namespace

extern "C" int lt_student_cnn_forward(const LtStudentCnnArgs* a, void* stream) {
  if (!a || a->M <= 0 || !a->out || (!a->image && !a->packed)) return LT_ERR_INVALID_ARG;
  if (!a->w1 || !a->b1 || !a->w2 || !a->b2 || !a->w3 || !a->b3 || !a->wh || !a->bh) return LT_ERR_INVALID_ARG;
  // geometry of the LocoTouch student pre-encoder (model_cfg.py:17-25); anything else stays with the torch modules
  if (a->in_channels != C0 || a->height != H0 || a->width != W0 || a->channels[0] != C1 || a->channels[1] != C2 || a->channels[2] != C3 ||
      a->kernel_sizes[0] != K1 || a->kernel_sizes[1] != K2 || a->kernel_sizes[2] != K3 || a->pool[0] != 2 || a->pool[1] != 1 || a->pool[2] != 1 ||
      a->embedding_dim <= 0 || a->embedding_dim > EMAX)
    return LT_ERR_UNSUPPORTED;
  if (a->packed && a->packed_words * 32 < H0 * W0) return LT_ERR_INVALID_ARG;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(student_cnn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) return LT_ERR_CUDA;
    attr_set = true;
  }
  CnnParams p;
  p.image = a->image; p.packed = a->packed; p.words = a->packed_words; p.M = a->M; p.E = a->embedding_dim;
  p.w1 = a->w1; p.b1 = a->b1; p.w2 = a->w2; p.b2 = a->b2; p.w3 = a->w3; p.b3 = a->b3; p.wh = a->wh; p.bh = a->bh; p.out = a->out;
  const int want = (int)lt::ceil_div(a->M, kWarps);
  const int grid = want < lt::sm_count() ? want : lt::sm_count();
  student_cnn_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(p);
  return lt::check_launch();
}
