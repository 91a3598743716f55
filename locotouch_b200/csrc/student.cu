// K8 helpers for the CNN-RNN student batch.
//  * lt_pad_trajectories -- ReplayBuffer._prepare_padded_sequence (reference locotouch/distill/replay_buffer.py:90-112):
//    the per-trajectory Python copy loop becomes one launch that writes the zero-padded [L_max, B, D] batch and its mask.
//  * lt_masked_mse       -- the behaviour-cloning loss of Student.train_on_data (reference locotouch/distill/student.py:
//    131 per-element MSE .mean(-1), :142 masked mean, :147-151 masked MAE) with its analytic gradient w.r.t. the
//    student actions; deterministic two-stage reduction, no host synchronisation.
#include "lt_common.cuh"

namespace {

constexpr int kThreads = 256;

// one warp per (t, b) output row
__global__ void __launch_bounds__(kThreads)
pad_trajectories_kernel(const float* __restrict__ flat, const int64_t* __restrict__ offsets, const int64_t* __restrict__ lengths,
                        int B, int L_max, int D, float* __restrict__ out, uint8_t* __restrict__ masks) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = (int64_t)L_max * B;
  const bool vec2 = (D & 1) == 0 && ((((uintptr_t)flat | (uintptr_t)out) & 7) == 0);
  for (int64_t r = warp; r < rows; r += nwarps) {
    const int t = (int)(r / B), b = (int)(r % B);
    const bool valid = t < lengths[b];
    if (masks && lane == 0) masks[r] = valid ? 1 : 0;
    float* dst = out + (size_t)r * D;
    const float* src = flat + (size_t)(offsets[b] + t) * D;
    if (vec2) {
      for (int i = lane; i < (D >> 1); i += 32)
        __stcs(reinterpret_cast<float2*>(dst) + i, valid ? __ldcs(reinterpret_cast<const float2*>(src) + i) : make_float2(0.f, 0.f));
    } else {
      for (int i = lane; i < D; i += 32) __stcs(dst + i, valid ? __ldcs(src + i) : 0.f);
    }
  }
}

struct MseWs {
  unsigned int counter;
  unsigned int pad[3];
  double partial[1];  // [blocks][3]
};

__global__ void __launch_bounds__(kThreads)
masked_mse_reduce_kernel(const float* __restrict__ s, const float* __restrict__ t, const uint8_t* __restrict__ masks, int64_t rows,
                         int A, float* __restrict__ out, MseWs* ws) {
  __shared__ double red[3][kThreads / 32];
  __shared__ bool is_last;
  double a_mse = 0.0, a_mae = 0.0, a_cnt = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < rows; r += (int64_t)gridDim.x * kThreads) {
    if (!masks[r]) continue;
    float se = 0.f, ae = 0.f;
    for (int j = 0; j < A; ++j) {
      const float d = s[r * A + j] - t[r * A + j];
      se += d * d;
      ae += fabsf(d);
    }
    a_mse += (double)(se / (float)A);
    a_mae += (double)(ae / (float)A);
    a_cnt += 1.0;
  }
  a_mse = lt::block_sum(a_mse, red[0]);
  a_mae = lt::block_sum(a_mae, red[1]);
  a_cnt = lt::block_sum(a_cnt, red[2]);
  if (threadIdx.x == 0) {
    ws->partial[3 * blockIdx.x] = a_mse;
    ws->partial[3 * blockIdx.x + 1] = a_mae;
    ws->partial[3 * blockIdx.x + 2] = a_cnt;
    __threadfence();
    is_last = atomicAdd(&ws->counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double x = 0.0, y = 0.0, z = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) {
    x += __ldcg(&ws->partial[3 * i]);
    y += __ldcg(&ws->partial[3 * i + 1]);
    z += __ldcg(&ws->partial[3 * i + 2]);
  }
  x = lt::block_sum(x, red[0]);
  y = lt::block_sum(y, red[1]);
  z = lt::block_sum(z, red[2]);
  if (threadIdx.x == 0) {
    out[0] = (float)(x / z);
    out[1] = (float)(y / z);
    out[2] = (float)z;
    out[3] = 0.f;
    ws->counter = 0;
  }
}

__global__ void __launch_bounds__(kThreads)
masked_mse_grad_kernel(const float* __restrict__ s, const float* __restrict__ t, const uint8_t* __restrict__ masks, int64_t rows,
                       int A, const float* __restrict__ out, float* __restrict__ grad) {
  const float scale = 2.0f / ((float)A * out[2]);
  const int64_t total = rows * A;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int64_t r = i / A;
    grad[i] = masks[r] ? (s[i] - t[i]) * scale : 0.f;
  }
}

}  // namespace

extern "C" int lt_pad_trajectories(const float* flat, const int64_t* offsets, const int64_t* lengths, int B, int L_max, int D,
                                   float* out, uint8_t* masks, void* stream) {
  if (!flat || !offsets || !lengths || !out || B <= 0 || L_max <= 0 || D <= 0) return LT_ERR_INVALID_ARG;
  int64_t blocks = lt::ceil_div((int64_t)L_max * B, kThreads / 32);
  const int64_t cap = 16LL * lt::sm_count();
  if (blocks > cap) blocks = cap;
  pad_trajectories_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(flat, offsets, lengths, B, L_max, D, out, masks);
  return lt::check_launch();
}

extern "C" int64_t lt_masked_mse_workspace_bytes(int64_t rows) {
  (void)rows;
  return 16 + 3 * 1024 * (int64_t)sizeof(double);
}

extern "C" int lt_masked_mse(const float* student, const float* teacher, const uint8_t* masks, int64_t rows, int A, float* grad_student,
                             float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!student || !teacher || !masks || !out || !workspace || rows <= 0 || A <= 0) return LT_ERR_INVALID_ARG;
  if (workspace_bytes < lt_masked_mse_workspace_bytes(rows)) return LT_ERR_WORKSPACE;
  int64_t blocks = lt::ceil_div(rows, kThreads);
  if (blocks > 1024) blocks = 1024;
  cudaStream_t st = (cudaStream_t)stream;
  masked_mse_reduce_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(student, teacher, masks, rows, A, out, (MseWs*)workspace);
  int rc = lt::check_launch();
  if (rc != LT_OK || !grad_student) return rc;
  int64_t gblocks = lt::ceil_div(rows * A, kThreads);
  const int64_t cap = 16LL * lt::sm_count();
  if (gblocks > cap) gblocks = cap;
  masked_mse_grad_kernel<<<(unsigned)gblocks, kThreads, 0, st>>>(student, teacher, masks, rows, A, out, grad_student);
  return lt::check_launch();
}
