// Shared device / host helpers for the locotouch_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "locotouch_b200.h"

#define LT_WARP 32
#define LT_FULL_MASK 0xffffffffu

namespace lt {

// ---------------------------------------------------------------------------------------------------- host side
void set_last_cuda_error(cudaError_t e);

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return LT_ERR_CUDA;
  }
  return LT_OK;
}

inline int check(cudaError_t e) {
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return LT_ERR_CUDA;
  }
  return LT_OK;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Number of SMs of the current device (148 on B200); cached.
int sm_count();

// ---------------------------------------------------------------------------------------------------- device side
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LT_FULL_MASK, v, o);
  return v;
}

template <typename T, int W>
__device__ __forceinline__ T group_sum(T v) {  // sum over aligned groups of W lanes
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(LT_FULL_MASK, v, o);
  return v;
}

// Block-wide sum in a fixed order (deterministic).  `smem` needs blockDim.x/32 elements.  Result valid in thread 0.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  T r = T(0);
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? smem[lane] : T(0);
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// Streaming (read-once) loads / stores: keep them out of L1.
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_stream4(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream4(float4* p, float4 v) { __stcs(p, v); }

// ------------------------------------------------------------------------------------------ Philox4x32-10 (counter RNG)
struct Philox {
  static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  __device__ static __forceinline__ uint4 round(uint4 c, uint2 k) {
    const uint32_t hi0 = __umulhi(kM0, c.x), lo0 = kM0 * c.x;
    const uint32_t hi1 = __umulhi(kM1, c.z), lo1 = kM1 * c.z;
    return make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
  }
  // 4 x 32 random bits for (seed, offset, a, b)
  __device__ static __forceinline__ uint4 gen(uint64_t seed, uint64_t offset, uint32_t a, uint32_t b) {
    uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 c = make_uint4(a, b, (uint32_t)offset, (uint32_t)(offset >> 32));
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      c = round(c, k);
      k.x += kW0;
      k.y += kW1;
    }
    return c;
  }
  // uniform in [0,1) with 24 random bits, the same construction torch uses for float32
  __device__ static __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }
  // two standard normals from two words (Box-Muller)
  __device__ static __forceinline__ float2 normal2(uint32_t x, uint32_t y) {
    const float u1 = ((x >> 8) + 1) * (1.0f / 16777216.0f);  // (0,1]
    const float u2 = (y >> 8) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(r * c, r * s);
  }
};

}  // namespace lt
