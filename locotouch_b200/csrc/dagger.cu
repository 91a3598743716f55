// K11: device-side DAgger replay buffer bookkeeping.
// Replaces the per-step host logic of reference locotouch/distill/replay_buffer.py:52-80 (collect_data: dones.any(),
// nonzero, .cpu().tolist() x2, a Python loop over done envs with .item() per env, per-trajectory torch.stack of per-step
// slices) and the packing half of :82-112.
//
//   lt_dagger_step        once per env step, one block: reward sums, episode log (reward, length) for every finished env, and --
//                         in env-index order, while the recorded-step budget lasts, exactly like the reference's
//                         `for env_id in done_idx: if steps_count - start_count < num_steps: record ... else: break` -- one
//                         (env, first step, length) record per finished trajectory.  The budget test over done envs in index
//                         order is an exclusive prefix sum of their lengths (block scan).
//   lt_pack_trajectories  after the collection: the recorded trajectories, which live as strided rows of the step-major
//                         [S, N, D] collection buffers, are copied back to back into the flat [total_steps, D] store the batch
//                         padding kernel (lt_pad_trajectories) reads; one warp per row, binary search of the row's trajectory.
#include "lt_common.cuh"

namespace {

constexpr int kScanThreads = 1024;

// state: [0] recorded steps (ReplayBuffer._steps_count), [1] trajectories recorded in this collection, [2] episodes logged in
// this collection, [3] unused
__global__ void __launch_bounds__(kScanThreads) dagger_step_kernel(const uint8_t* __restrict__ dones, const float* __restrict__ reward,
                                                                  float* __restrict__ reward_sums, int32_t* __restrict__ start_idx, int N,
                                                                  int step_now, long long limit, int always_restart, long long* __restrict__ state,
                                                                  int32_t* __restrict__ traj_env, int32_t* __restrict__ traj_start,
                                                                  int32_t* __restrict__ traj_len, float* __restrict__ ep_reward,
                                                                  int32_t* __restrict__ ep_length) {
  __shared__ long long s_len[kScanThreads / 32];
  __shared__ int s_cnt[kScanThreads / 32];
  __shared__ long long s_rec_len[kScanThreads / 32];
  __shared__ int s_rec_cnt[kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = (N + kScanThreads - 1) / kScanThreads;
  const int n0 = tid * chunk, n1 = min(N, n0 + chunk);
  // pass 1: per-thread totals over its consecutive envs
  long long my_len = 0;
  int my_cnt = 0;
  for (int n = n0; n < n1; ++n) {
    if (dones[n]) {
      my_len += step_now - start_idx[n];
      ++my_cnt;
    }
  }
  // block exclusive scan of (length, count) over threads
  long long inc_len = my_len;
  int inc_cnt = my_cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long l = __shfl_up_sync(LT_FULL_MASK, inc_len, o);
    const int c = __shfl_up_sync(LT_FULL_MASK, inc_cnt, o);
    if (lane >= o) { inc_len += l; inc_cnt += c; }
  }
  if (lane == 31) { s_len[warp] = inc_len; s_cnt[warp] = inc_cnt; }
  __syncthreads();
  long long pre_len = inc_len - my_len;
  int pre_cnt = inc_cnt - my_cnt;
  for (int w = 0; w < warp; ++w) { pre_len += s_len[w]; pre_cnt += s_cnt[w]; }
  const long long count0 = state[0];
  const int traj0 = (int)state[1], ep0 = (int)state[2];
  // pass 2: in env order
  long long running = count0 + pre_len, rec_len = 0;
  int k = pre_cnt, rec_cnt = 0;
  for (int n = n0; n < n1; ++n) {
    const float total = reward_sums[n] + reward[n];  // replay_buffer.py:57: accumulated before the done test
    if (!dones[n]) {
      reward_sums[n] = total;
      continue;
    }
    const int len = step_now - start_idx[n];
    ep_reward[ep0 + k] = total;  // :64-65 (every finished env, recorded or not)
    ep_length[ep0 + k] = len;
    reward_sums[n] = 0.f;        // :66
    if (running < limit) {       // :68 budget test, in env-index order; once it fails it fails for every later env (:73 break)
      traj_env[traj0 + k] = n;
      traj_start[traj0 + k] = start_idx[n];
      traj_len[traj0 + k] = len;
      start_idx[n] = step_now;   // :71
      rec_len += len;
      ++rec_cnt;
    } else if (always_restart) {
      start_idx[n] = step_now;
    }
    running += len;
    ++k;
  }
  // totals -> state (every thread has finished reading state above: the barrier below orders the update)
  rec_len = lt::warp_sum(rec_len);
  rec_cnt = lt::warp_sum(rec_cnt);
  if (lane == 0) { s_rec_len[warp] = rec_len; s_rec_cnt[warp] = rec_cnt; }
  __syncthreads();
  if (tid == 0) {
    long long tl = 0;
    int tc = 0, done_total = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) { tl += s_rec_len[w]; tc += s_rec_cnt[w]; done_total += s_cnt[w]; }
    state[0] = count0 + tl;
    state[1] = traj0 + tc;
    state[2] = ep0 + done_total;
  }
}

__global__ void __launch_bounds__(256) pack_trajectories_kernel(const float* __restrict__ x, const int32_t* __restrict__ traj_env,
                                                               const int32_t* __restrict__ traj_start, const int64_t* __restrict__ traj_offset,
                                                               int M, int64_t total_rows, int N, int D, float* __restrict__ flat, int vec) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < total_rows; r += (int64_t)gridDim.x * 8) {
    int lo = 0, hi = M - 1;  // last trajectory whose offset <= r
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (traj_offset[mid] <= r) lo = mid; else hi = mid - 1;
    }
    const int64_t pos = r - traj_offset[lo];
    const float* src = x + ((size_t)(traj_start[lo] + pos) * N + traj_env[lo]) * D;
    float* dst = flat + (size_t)r * D;
    if (vec) {
      for (int i = lane; i < (D >> 2); i += 32) reinterpret_cast<float4*>(dst)[i] = __ldcs(reinterpret_cast<const float4*>(src) + i);
    } else {
      for (int i = lane; i < D; i += 32) dst[i] = __ldcs(src + i);
    }
  }
}

}  // namespace

extern "C" int lt_dagger_step(const uint8_t* dones, const float* reward, float* reward_sums, int32_t* start_idx, int N, int step_now,
                              int64_t limit, int always_restart, int64_t* state, int32_t* traj_env, int32_t* traj_start, int32_t* traj_len,
                              float* ep_reward, int32_t* ep_length, void* stream) {
  if (!dones || !reward || !reward_sums || !start_idx || !state || !traj_env || !traj_start || !traj_len || !ep_reward || !ep_length)
    return LT_ERR_INVALID_ARG;
  if (N <= 0 || N > kScanThreads * 64) return LT_ERR_INVALID_ARG;
  dagger_step_kernel<<<1, kScanThreads, 0, (cudaStream_t)stream>>>(dones, reward, reward_sums, start_idx, N, step_now, (long long)limit,
                                                                  always_restart, reinterpret_cast<long long*>(state), traj_env, traj_start,
                                                                  traj_len, ep_reward, ep_length);
  return lt::check_launch();
}

extern "C" int lt_pack_trajectories(const float* x, const int32_t* traj_env, const int32_t* traj_start, const int64_t* traj_offset, int M,
                                    int64_t total_rows, int N, int D, float* flat, void* stream) {
  if (!x || !traj_env || !traj_start || !traj_offset || !flat || M <= 0 || total_rows <= 0 || N <= 0 || D <= 0) return LT_ERR_INVALID_ARG;
  const int vec = (D % 4 == 0) && (((uintptr_t)x | (uintptr_t)flat) & 15) == 0;
  int64_t blocks = lt::ceil_div(total_rows, 8);
  const int64_t cap = (int64_t)lt::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  pack_trajectories_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, traj_env, traj_start, traj_offset, M, total_rows, N, D, flat, vec);
  return lt::check_launch();
}
