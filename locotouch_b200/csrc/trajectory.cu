// K10: trajectory split / pad / unpad for recurrent mini-batches.
// Replaces reference loco_rl/loco_rl/utils/utils.py:37-83 (split_and_pad_trajectories: clone + nonzero + tolist + torch.split
// into thousands of views + pad_sequence; unpad_trajectories: boolean-mask gather) as used by
// RolloutStorage.recurrent_mini_batch_generator (storage/rollout_storage.py:246-318).
//
// A rollout [T, N, D] is cut at every done (the last step always ends a trajectory); trajectories are numbered env by env in
// time order (the reference's transpose(1,0).flatten order).  traj_base[n] = number of trajectories of the envs before n
// (an exclusive scan of 1 + #dones before the last step, done by the caller, who also needs the total M to size the output).
//   lt_trajectory_index     one thread per env walks its T dones once: (env, start, length) of each of its trajectories
//   lt_split_pad_...        OUTPUT-driven: one warp per (position, trajectory) row of out[T, M, D] copies the source row or
//                           writes zeros -> coalesced 16-byte stores, no memset pass, no atomics; masks[T, M] from the lengths
//   lt_unpad_...            the inverse: every valid (position, trajectory) row goes back to out[start + position, env, :]
#include "lt_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void trajectory_index_kernel(const uint8_t* __restrict__ dones, const int64_t* __restrict__ traj_base, int32_t* __restrict__ traj_env,
                                        int32_t* __restrict__ traj_start, int32_t* __restrict__ traj_len, int T, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int traj = (int)traj_base[n], start = 0;
  for (int t = 0; t < T; ++t) {
    if (t == T - 1 || dones[(size_t)t * N + n]) {  // utils.py:55-56: dones[-1] = 1
      traj_env[traj] = n;
      traj_start[traj] = start;
      traj_len[traj] = t + 1 - start;
      ++traj;
      start = t + 1;
    }
  }
}

// rows = T * M (position-major: row = pos * M + traj).  SCATTER == false: out[pos, traj] = valid ? x[start + pos, env] : 0.
// SCATTER == true: x is [T, M, D] padded, out is [T, N, D]: out[start + pos, env] = x[pos, traj] for valid rows.
template <bool SCATTER>
__global__ void __launch_bounds__(kThreads) trajectory_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ traj_env,
                                                                  const int32_t* __restrict__ traj_start, const int32_t* __restrict__ traj_len,
                                                                  float* __restrict__ out, uint8_t* __restrict__ masks, int T, int N, int D, int M,
                                                                  int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)T * M;
  for (int64_t row = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * (kThreads / 32)) {
    const int pos = (int)(row / M), traj = (int)(row - (int64_t)pos * M);
    const int len = traj_len[traj];
    const bool valid = pos < len;
    if (masks && lane == 0) masks[row] = valid;
    const size_t flat = ((size_t)(traj_start[traj] + pos) * N + traj_env[traj]) * D;  // row of the [T, N, D] tensor
    const size_t padded = (size_t)row * D;
    if (SCATTER) {
      if (!valid) continue;
      if (vec) {
        for (int i = lane; i < (D >> 2); i += 32) reinterpret_cast<float4*>(out + flat)[i] = __ldcs(reinterpret_cast<const float4*>(x + padded) + i);
      } else {
        for (int i = lane; i < D; i += 32) out[flat + i] = __ldcs(x + padded + i);
      }
    } else {
      if (vec) {
        for (int i = lane; i < (D >> 2); i += 32) {
          const float4 v = valid ? __ldcs(reinterpret_cast<const float4*>(x + flat) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          __stcs(reinterpret_cast<float4*>(out + padded) + i, v);
        }
      } else {
        for (int i = lane; i < D; i += 32) __stcs(out + padded + i, valid ? __ldcs(x + flat + i) : 0.f);
      }
    }
  }
}

int launch_rows(bool scatter, const float* x, const int32_t* env, const int32_t* start, const int32_t* len, float* out, uint8_t* masks, int T, int N,
                int D, int M, void* stream) {
  if (!x || !env || !start || !len || !out || T <= 0 || N <= 0 || D <= 0 || M <= 0) return LT_ERR_INVALID_ARG;
  const int vec = (D % 4 == 0) && (((uintptr_t)x | (uintptr_t)out) & 15) == 0;
  const int64_t rows = (int64_t)T * M;
  int64_t blocks = lt::ceil_div(rows, kThreads / 32);
  const int64_t cap = (int64_t)lt::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (scatter)
    trajectory_rows_kernel<true><<<(unsigned)blocks, kThreads, 0, st>>>(x, env, start, len, out, masks, T, N, D, M, vec);
  else
    trajectory_rows_kernel<false><<<(unsigned)blocks, kThreads, 0, st>>>(x, env, start, len, out, masks, T, N, D, M, vec);
  return lt::check_launch();
}

}  // namespace

extern "C" int lt_trajectory_index(const uint8_t* dones, const int64_t* traj_base, int32_t* traj_env, int32_t* traj_start, int32_t* traj_len, int T,
                                   int N, void* stream) {
  if (!dones || !traj_base || !traj_env || !traj_start || !traj_len || T <= 0 || N <= 0) return LT_ERR_INVALID_ARG;
  trajectory_index_kernel<<<(unsigned)lt::ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(dones, traj_base, traj_env, traj_start, traj_len, T, N);
  return lt::check_launch();
}

extern "C" int lt_split_pad_trajectories(const float* x, const int32_t* traj_env, const int32_t* traj_start, const int32_t* traj_len, float* out,
                                         uint8_t* masks, int T, int N, int D, int M, void* stream) {
  return launch_rows(false, x, traj_env, traj_start, traj_len, out, masks, T, N, D, M, stream);
}

extern "C" int lt_unpad_trajectories(const float* padded, const int32_t* traj_env, const int32_t* traj_start, const int32_t* traj_len, float* out,
                                     int T, int N, int D, int M, void* stream) {
  return launch_rows(true, padded, traj_env, traj_start, traj_len, out, nullptr, T, N, D, M, stream);
}
