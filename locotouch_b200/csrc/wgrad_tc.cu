// K15: weight (and bias) gradient of one Linear layer of the actor-critic / student MLPs,
//     dW[n, k] (+)= sum_b g[b, n] * x[b, k],   db[n] (+)= sum_b g[b, n]      (B = 24576 rows of a PPO mini-batch, n x k <= 512 x 512)
// as ONE hand-written tcgen05 kernel with in-kernel split-K: replaces, per nn.Linear of reference
// loco_rl/loco_rl/modules/actor_critic.py:33-56 (autograd's `grad_output.t().mm(input)` and `grad_output.sum(0)` in
// algorithms/ppo.py:350), the 8-way torch.bmm split + torch.sum (cuBLAS tf32gemm / an sm_80 s1688gemm / splitKreduce / ATen
// reduce_kernel launches) and the separate bias-gradient reduction pass (K9) of round 1.
//
// GEMM view: M = n (columns of g), N = k (columns of x), K = B (the batch).  Both operands are consumed exactly as they lie in HBM:
// g [B, n] and x [B, k] are row-major, i.e. "MN-major" operands for the tensor core -- tcgen05.mma kind::tf32 takes MN-major A and B
// from shared memory (wgmma could not for 32-bit types), so no transpose pass exists anywhere.
//
// What bounds it: fp32 operands make the shared-memory fill the scarce resource (a 128 x 256 tile needs 96 B/clk per SM at the
// TF32 MMA rate; the L2 delivers ~42 B/clk per SM with all SMs pulling).  So one CTA owns a 128-row slab of dW over its FULL width
// (up to 512 columns = the whole TMEM: NH accumulators of 128 x BNH) and a slice of the batch: a batch row costs (128 + k) * 4
// bytes of fill instead of (128 + 256) * 4 per 256 columns, and g is read n/128 times in total instead of (n/128) * (k/256).
//
// One CTA: warp 0 streams [16 x 128] / [16 x BNH] boxes of g and x into a STAGES-deep shared-memory ring with TMA (mbarrier
// expect-tx), warp 1 issues 2 * NH tcgen05.mma (K = 8 each) per stage into the TMEM accumulators and releases the stage with
// tcgen05.commit; warps 2-5 meanwhile add up the rows of g that pass through shared memory (thread = one of the 128 gradient
// columns: the bias gradient costs no extra byte of traffic), and when the slice is done read the accumulators back
// (tcgen05.ld 32x32b) and add them into the flat gradient buffer with 16-byte vector reductions (red.global.add.v4.f32), each
// CTA starting at a different column block so that the splits do not queue up on the same L2 lines -- the split-K partials never
// exist in memory.  Grid = row slabs x batch slices sized to one CTA per SM.
// Layouts / descriptors come from CuTe (UMMA::Layout_MN_SW128_32B_Atom -- the one shared-memory layout tcgen05 accepts for
// MN-major 32-bit operands --, make_tma_atom, make_umma_desc through the MMA atom); the kernel, its pipeline, the split-K
// schedule and the epilogue are ours.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "lt_common.cuh"

#if LT_HAVE_CUTLASS

#include <cute/tensor.hpp>
#include <cute/arch/tmem_allocator_sm100.hpp>
#include <cute/atom/mma_atom.hpp>
#include <cute/atom/copy_traits_sm90_tma.hpp>
#include <cutlass/arch/barrier.h>

namespace lt_wgrad {

using namespace cute;
using TF = cute::tfloat32_t;

constexpr int kBM = 128;   // rows of the output slab = one UMMA M
constexpr int kBK = 16;    // batch rows per pipeline stage = 2 UMMA K steps
constexpr int kThreads = 192;

template <int BNH, int NH>
struct Cfg {
  static constexpr int kNT = BNH * NH;                         // dW columns per CTA
  static constexpr int kTmemCols = kNT <= 32 ? 32 : kNT <= 64 ? 64 : kNT <= 128 ? 128 : kNT <= 256 ? 256 : 512;
  static constexpr int kStageBytes = (kBM + kNT) * kBK * 4;
  static constexpr int kStages = (200 * 1024 / kStageBytes) < 10 ? (200 * 1024 / kStageBytes) : 10;
  using Mma = decltype(make_tiled_mma(SM100_MMA_TF32_SS<TF, TF, float, kBM, BNH, UMMA::Major::MN, UMMA::Major::MN>{}));
  using ShapeA = decltype(partition_shape_A(Mma{}, make_shape(Int<kBM>{}, Int<kBK>{})));
  using ShapeB = decltype(partition_shape_B(Mma{}, make_shape(Int<BNH>{}, Int<kBK>{})));
  // ((MMA_MN, MMA_K), MNs, Ks, slot); K-blocks of one slot are laid out first, like the CUTLASS collectives do for MN-major operands.
  // B has NH slots per stage (slot = stage * NH + half).
  using SmemA = decltype(UMMA::tile_to_mma_shape(UMMA::Layout_MN_SW128_32B_Atom<TF>{}, append(ShapeA{}, Int<kStages>{}), Step<_2, _1, _3>{}));
  using SmemB = decltype(UMMA::tile_to_mma_shape(UMMA::Layout_MN_SW128_32B_Atom<TF>{}, append(ShapeB{}, Int<kStages * NH>{}), Step<_2, _1, _3>{}));
  struct Storage {
    alignas(1024) cute::ArrayEngine<TF, cute::cosize_v<SmemA>> a;
    alignas(1024) cute::ArrayEngine<TF, cute::cosize_v<SmemB>> b;
    alignas(16) uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t acc_full;
    uint32_t tmem_base;
  };
  static constexpr int kSmemBytes = (int)sizeof(Storage) + 1024;  // + slack for the manual 1024-byte alignment
  // epilogue staging: [128 x 32] fp32 boxes (16 KB, rows of 128 bytes, 128B-swizzled) that alias the operand ring once the last
  // MMA has read it; each box leaves through ONE bulk tensor reduction (cp.reduce.async.bulk.tensor ... .add)
  static constexpr int kBlocks = kNT / 32;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kOutBufs = (kRingBytes / 16384) < kBlocks ? (kRingBytes / 16384) : kBlocks;
  using SmemD = decltype(tile_to_shape(UMMA::Layout_K_SW128_Atom<float>{}, make_shape(Int<kBM>{}, Int<32>{}, Int<kOutBufs>{})));
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t& bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cute::cast_smem_ptr_to_uint(&bar)) : "memory");
}

template <int BNH, int NH, class TmaA, class TmaB, class TmaD, class TensorA, class TensorB, class TensorD>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_splitk_kernel(TensorA mA, TensorB mB, TensorD mD, float* __restrict__ dw_0, float* __restrict__ dbias_0, int n_out, int k_in, int k_tiles,
                    int k_tiles_per_split, int bulk_reduce, CUTE_GRID_CONSTANT TmaA const tma_a0, CUTE_GRID_CONSTANT TmaB const tma_b0,
                    CUTE_GRID_CONSTANT TmaD const tma_d0, int splits, float* __restrict__ dw_1, float* __restrict__ dbias_1,
                    CUTE_GRID_CONSTANT TmaA const tma_a1, CUTE_GRID_CONSTANT TmaB const tma_b1, CUTE_GRID_CONSTANT TmaD const tma_d1) {
  using C = Cfg<BNH, NH>;
  constexpr int S = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  typename C::Storage& ss = *reinterpret_cast<typename C::Storage*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  // two problems of the same shape (the actor's and the critic's layer) share one launch: blockIdx.z >= splits works on the second
  const bool second = (int)blockIdx.z >= splits;
  const int zsplit = (int)blockIdx.z - (second ? splits : 0);
  TmaA const& tma_a = second ? tma_a1 : tma_a0;
  TmaB const& tma_b = second ? tma_b1 : tma_b0;
  TmaD const& tma_d = second ? tma_d1 : tma_d0;
  float* __restrict__ dw = second ? dw_1 : dw_0;
  float* __restrict__ dbias = second ? dbias_1 : dbias_0;

  const int warp = threadIdx.x >> 5;
  const int kt0 = zsplit * k_tiles_per_split;
  const int nkt = min(k_tiles, kt0 + k_tiles_per_split) - kt0;
  if (nkt <= 0) return;  // uniform per CTA (only when splits * per_split overshoots)
  const bool fold_bias = dbias != nullptr && blockIdx.y == 0;  // uniform per CTA

  typename C::Mma tiled_mma;
  Tensor gA = local_tile(mA, make_shape(Int<kBM>{}, Int<kBK>{}), make_coord(blockIdx.x, _));           // (BM, BK, k_tiles)
  Tensor gB = local_tile(mB, make_shape(Int<BNH>{}, Int<kBK>{}), make_coord(_, _));                    // (BNH, BK, halves, k_tiles)
  Tensor sA = make_tensor(make_smem_ptr(ss.a.begin()), typename C::SmemA{});
  Tensor sB = make_tensor(make_smem_ptr(ss.b.begin()), typename C::SmemB{});
  ThrMMA cta_mma = tiled_mma.get_slice(0);
  Tensor tCgA = cta_mma.partition_A(gA);      // ((MMA_M, MMA_K), Ms, Ks, k_tiles)
  Tensor tCgB = cta_mma.partition_B(gB);      // ((MMA_N, MMA_K), Ns, Ks, halves, k_tiles)
  Tensor tCrA = cta_mma.make_fragment_A(sA);  // shared-memory matrix descriptors, (1, MNs, Ks, slot)
  Tensor tCrB = cta_mma.make_fragment_B(sB);
  Tensor tCtAcc = tiled_mma.make_fragment_C(partition_shape_C(tiled_mma, make_shape(Int<kBM>{}, Int<BNH>{})));

  auto [tAgA, tAsA] = tma_partition(tma_a, Int<0>{}, Layout<_1>{}, group_modes<0, 3>(sA), group_modes<0, 3>(tCgA));
  auto [tBgB, tBsB] = tma_partition(tma_b, Int<0>{}, Layout<_1>{}, group_modes<0, 3>(sB), group_modes<0, 3>(tCgB));

  if (warp == 0 && cute::elect_one_sync()) {
    cute::prefetch_tma_descriptor(tma_a.get_tma_descriptor());
    cute::prefetch_tma_descriptor(tma_b.get_tma_descriptor());
    cute::prefetch_tma_descriptor(tma_d.get_tma_descriptor());
    for (int s = 0; s < S; ++s) {
      cute::initialize_barrier(ss.full[s], 1);
      cute::initialize_barrier(ss.empty[s], fold_bias ? 5 : 1);  // tcgen05.commit (+ one arrival per bias-folding warp)
    }
    cute::initialize_barrier(ss.acc_full, 1);
    cutlass::arch::fence_barrier_init();
  }
  cute::TMEM::Allocator1Sm tmem;
  if (warp == 1) {
    tmem.allocate(C::kTmemCols, &ss.tmem_base);
    tmem.release_allocation_lock();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = ss.tmem_base;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (cute::elect_one_sync()) {
      for (int i = 0; i < nkt; ++i) {
        const int s = i % S;
        if (i >= S) cute::wait_barrier(ss.empty[s], ((i / S) - 1) & 1);
        cute::set_barrier_transaction_bytes(ss.full[s], C::kStageBytes);
        copy(tma_a.with(ss.full[s]), tAgA(_, kt0 + i), tAsA(_, s));
        CUTE_UNROLL
        for (int h = 0; h < NH; ++h) copy(tma_b.with(ss.full[s]), tBgB(_, blockIdx.y * NH + h, kt0 + i), tBsB(_, s * NH + h));
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (the atom elects one lane itself)
    uint32_t accumulate = 0;
    for (int i = 0; i < nkt; ++i) {
      const int s = i % S;
      cute::wait_barrier(ss.full[s], (i / S) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      CUTE_UNROLL
      for (int h = 0; h < NH; ++h) {
        tCtAcc.data() = tmem_base + h * BNH;
        tiled_mma.accumulate_ = accumulate ? UMMA::ScaleOut::One : UMMA::ScaleOut::Zero;
        CUTE_UNROLL
        for (int kb = 0; kb < size<2>(tCrA); ++kb) {
          gemm(tiled_mma, tCrA(_, _, kb, s), tCrB(_, _, kb, s * NH + h), tCtAcc);
          tiled_mma.accumulate_ = UMMA::ScaleOut::One;
        }
      }
      accumulate = 1;
      cutlass::arch::umma_arrive(&ss.empty[s]);  // tcgen05.commit: the stage is free once these MMAs have read it
    }
    cutlass::arch::umma_arrive(&ss.acc_full);
  } else {
    // ---------------------------------------------------------------- bias fold, then epilogue: TMEM -> registers -> red.add
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int m_local = q * 32 + (threadIdx.x & 31);
    const int row = blockIdx.x * kBM + m_local;
    if (fold_bias) {
      float bsum = 0.0f;
      for (int i = 0; i < nkt; ++i) {
        const int s = i % S;
        cute::wait_barrier(ss.full[s], (i / S) & 1);
        asm volatile("" ::: "memory");  // cute::wait_barrier carries no memory clobber: keep the plain shared-memory loads below it
        CUTE_UNROLL
        for (int kb = 0; kb < size<2>(sA); ++kb) {
          CUTE_UNROLL
          for (int k8 = 0; k8 < 8; ++k8) {  // explicit ld.shared: a generic LD.E is not ordered with the mbarrier unit at all
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(cute::cast_smem_ptr_to_uint(raw_pointer_cast(&sA(make_coord(m_local, k8), 0, kb, s)))) : "memory");
            bsum += v;
          }
        }
        // The loads above are asynchronous and nothing in the arrive below waits for them: without a consumer in front of it the stage is
        // released -- and refilled by TMA -- while rows of it are still being fetched (seen as a few rows of the NEXT batch slice in the
        // column sums whenever TMA is fast: aligned row pitches, warm L2).  The sum is made a control dependency of the release.
        if (__float_as_uint(bsum) == 0x7fc0beefu) asm volatile("trap;");  // a real consumer of every load in front of the release (never taken)
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(ss.empty[s]);
      }
      if (row < n_out) atomicAdd(dbias + row, bsum);
    }
    cute::wait_barrier(ss.acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // the staging boxes below alias the operand ring: the MMAs are done with it (acc_full), but a slower bias-folding warp may still be
    // adding up the g rows of the last stages -- all four warps must have left the fold loop before the first box is written
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int col0 = blockIdx.y * C::kNT;
    float* out = dw + (size_t)row * k_in + col0;
    constexpr int kBlocks = C::kBlocks;
    constexpr int NB = C::kOutBufs;
    // staging boxes over the (now idle) operand ring; box cc % NB carries column block cc
    Tensor sD = make_tensor(make_smem_ptr(reinterpret_cast<float*>(ss.a.begin())), typename C::SmemD{});                 // (128, 32, NB)
    Tensor gD = local_tile(mD, make_shape(Int<kBM>{}, Int<32>{}), make_coord(blockIdx.x, _));                             // (128, 32, column blocks)
    auto [tDgD, tDsD] = tma_partition(tma_d, Int<0>{}, Layout<_1>{}, group_modes<0, 2>(sD), group_modes<0, 2>(gD));
    const bool issuer = threadIdx.x == 64;       // first lane of the first epilogue warp
    int issued = 0;
#pragma unroll 1
    for (int cc = 0; cc < kBlocks; ++cc) {
      const int cb = (cc + zsplit) % kBlocks;  // staggered start: the splits of one slab hit different lines
      const int c = cb * 32;
      if (col0 + c >= k_in) continue;            // uniform over the CTA
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (bulk_reduce) {
        const int buf = issued % NB;
        if (issued >= NB) {                      // the box is being reused: its previous reduction must have read it
          if (issuer) cute::tma_store_wait<NB - 1>();
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(&sD(m_local, j, buf)) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        cute::tma_store_fence();                 // generic-proxy writes -> visible to the async proxy
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          copy(tma_d, tDsD(_, buf), tDgD(_, col0 / 32 + cb));  // rows >= n_out / columns >= k_in are clipped by the tensor map
          cute::tma_store_arrive();
        }
        ++issued;
      } else if (row < n_out) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (col0 + c + j < k_in)  // k_in % 4 == 0: a float4 is inside or outside as a whole
            red_add_v4(out + c + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
    if (bulk_reduce && issuer) cute::tma_store_wait<0>();  // shared memory must outlive the reads of the bulk reductions
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) tmem.free(tmem_base, C::kTmemCols);
}

template <int BNH, int NH>
int launch(const float* g, const float* x, float* dw, float* dbias, int B, int n_out, int k_in, int sms, int bulk_reduce, cudaStream_t st,
           const float* g1 = nullptr, const float* x1 = nullptr, float* dw1 = nullptr, float* dbias1 = nullptr) {
  using C = Cfg<BNH, NH>;
  const int nprob = g1 ? 2 : 1;
  static_assert(C::kNT % 32 == 0 && C::kNT <= 512 && C::kStages >= 3, "tile configuration");
  // (MN, K) views with the MN mode contiguous: exactly the row-major [B, n] / [B, k] tensors
  Tensor mA = make_tensor(make_gmem_ptr(reinterpret_cast<TF const*>(g)), make_layout(make_shape(n_out, B), make_stride(Int<1>{}, n_out)));
  Tensor mB = make_tensor(make_gmem_ptr(reinterpret_cast<TF const*>(x)), make_layout(make_shape(k_in, B), make_stride(Int<1>{}, k_in)));
  typename C::SmemA sa;
  typename C::SmemB sb;
  auto tma_a = make_tma_atom(SM90_TMA_LOAD{}, mA, sa(_, _, _, Int<0>{}), make_shape(Int<kBM>{}, Int<kBK>{}));
  auto tma_b = make_tma_atom(SM90_TMA_LOAD{}, mB, sb(_, _, _, Int<0>{}), make_shape(Int<BNH>{}, Int<kBK>{}));
  Tensor cA = tma_a.get_tma_tensor(shape(mA));
  Tensor cB = tma_b.get_tma_tensor(shape(mB));
  Tensor mD = make_tensor(make_gmem_ptr(dw), make_layout(make_shape(n_out, k_in), make_stride(k_in, Int<1>{})));
  typename C::SmemD sd;
  auto tma_d = make_tma_atom(SM90_TMA_REDUCE_ADD{}, mD, sd(_, _, Int<0>{}), make_shape(Int<kBM>{}, Int<32>{}));
  Tensor cD = tma_d.get_tma_tensor(shape(mD));
  const int tiles_m = (n_out + kBM - 1) / kBM, tiles_n = (k_in + C::kNT - 1) / C::kNT;
  const int k_tiles = (B + kBK - 1) / kBK;
  int splits = sms / (tiles_m * tiles_n * nprob);
  if (splits < 1) splits = 1;
  if (splits > k_tiles) splits = k_tiles;
  const int per = (k_tiles + splits - 1) / splits;
  splits = (k_tiles + per - 1) / per;
  auto* kern = &wgrad_splitk_kernel<BNH, NH, decltype(tma_a), decltype(tma_b), decltype(tma_d), decltype(cA), decltype(cB), decltype(cD)>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) != cudaSuccess) return LT_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid(tiles_m, tiles_n, splits * nprob);
  if (nprob == 2) {
    Tensor mA1 = make_tensor(make_gmem_ptr(reinterpret_cast<TF const*>(g1)), make_layout(make_shape(n_out, B), make_stride(Int<1>{}, n_out)));
    Tensor mB1 = make_tensor(make_gmem_ptr(reinterpret_cast<TF const*>(x1)), make_layout(make_shape(k_in, B), make_stride(Int<1>{}, k_in)));
    Tensor mD1 = make_tensor(make_gmem_ptr(dw1), make_layout(make_shape(n_out, k_in), make_stride(k_in, Int<1>{})));
    auto tma_a1 = make_tma_atom(SM90_TMA_LOAD{}, mA1, sa(_, _, _, Int<0>{}), make_shape(Int<kBM>{}, Int<kBK>{}));
    auto tma_b1 = make_tma_atom(SM90_TMA_LOAD{}, mB1, sb(_, _, _, Int<0>{}), make_shape(Int<BNH>{}, Int<kBK>{}));
    auto tma_d1 = make_tma_atom(SM90_TMA_REDUCE_ADD{}, mD1, sd(_, _, Int<0>{}), make_shape(Int<kBM>{}, Int<32>{}));
    kern<<<grid, kThreads, C::kSmemBytes, st>>>(cA, cB, cD, dw, dbias, n_out, k_in, k_tiles, per, bulk_reduce, tma_a, tma_b, tma_d, splits, dw1, dbias1, tma_a1, tma_b1,
                                                tma_d1);
  } else {
    kern<<<grid, kThreads, C::kSmemBytes, st>>>(cA, cB, cD, dw, dbias, n_out, k_in, k_tiles, per, bulk_reduce, tma_a, tma_b, tma_d, splits, dw, dbias, tma_a, tma_b, tma_d);
  }
  return lt::check_launch();
}

// narrow output layers (the action-mean head n = 12, the value head n = 1): CUDA cores, ONE pass over x with 16-byte loads.
// A thread owns four consecutive columns of x and every (blockDim / quads)-th row of the block's row range; the g row it needs
// (<= 16 floats) is a warp-broadcast load.  Row groups are folded through shared memory, one vector reduction per block and quad.
template <int NMAX>
__global__ void __launch_bounds__(256, 2)
wgrad_narrow_kernel(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ dw, float* __restrict__ dbias, int B, int n_out, int k_in,
                    int rows_per_block) {
  __shared__ float4 fold[256];
  const int quads = k_in >> 2;                       // k_in % 4 == 0, quads <= 256 (checked by the caller)
  const int groups = blockDim.x / quads;             // rows in flight per block
  const int cq = threadIdx.x % quads, rg = threadIdx.x / quads;
  const bool live = rg < groups;
  const int b0 = blockIdx.x * rows_per_block;
  const int b1 = min(B, b0 + rows_per_block);
  float4 acc[NMAX];
  float bacc[NMAX];  // bias gradient = column sums of g: the first quad of every row group adds up the g values it loads anyway
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    bacc[j] = 0.0f;
  }
  const bool fold_bias = dbias != nullptr && cq == 0;
  if (live) {
#pragma unroll 4
    for (int r = b0 + rg; r < b1; r += groups) {
      const float4 xv = __ldcs(reinterpret_cast<const float4*>(x + (size_t)r * k_in) + cq);
      const float* gr = g + (size_t)r * n_out;
#pragma unroll
      for (int j = 0; j < NMAX; ++j)
        if (j < n_out) {
          const float gv = __ldg(gr + j);
          acc[j].x = fmaf(gv, xv.x, acc[j].x);
          acc[j].y = fmaf(gv, xv.y, acc[j].y);
          acc[j].z = fmaf(gv, xv.z, acc[j].z);
          acc[j].w = fmaf(gv, xv.w, acc[j].w);
          if (fold_bias) bacc[j] += gv;
        }
    }
  }
  if (dbias != nullptr) {  // fold the row groups of the block in shared memory: one atomic per block and output
    __shared__ float bfold[64][NMAX];  // groups <= 256 / 4
    if (live && cq == 0) {
#pragma unroll
      for (int j = 0; j < NMAX; ++j) bfold[rg][j] = bacc[j];
    }
    __syncthreads();
    if ((int)threadIdx.x < n_out) {
      float t = 0.0f;
      for (int q = 0; q < groups; ++q) t += bfold[q][threadIdx.x];
      atomicAdd(dbias + threadIdx.x, t);
    }
  }
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    if (j >= n_out) break;  // uniform
    __syncthreads();
    fold[threadIdx.x] = acc[j];
    __syncthreads();
    if (rg == 0) {
      float4 s = fold[cq];
      for (int q = 1; q < groups; ++q) {
        const float4 v = fold[q * quads + cq];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      red_add_v4(dw + (size_t)j * k_in + 4 * cq, s.x, s.y, s.z, s.w);
    }
  }
}

// any other narrow shape (k_in % 4 != 0 or k_in > 1024): one thread per column, scalar loads
__global__ void __launch_bounds__(256)
wgrad_narrow_scalar_kernel(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ dw, float* __restrict__ dbias, int B, int n_out, int k_in,
                           int rows_per_block) {
  const int b0 = blockIdx.x * rows_per_block;
  const int b1 = min(B, b0 + rows_per_block);
  if (dbias != nullptr && threadIdx.x < n_out) {
    float t = 0.0f;
    for (int r = b0; r < b1; ++r) t += __ldg(g + (size_t)r * n_out + threadIdx.x);
    atomicAdd(dbias + threadIdx.x, t);
  }
  for (int kc = threadIdx.x; kc < k_in; kc += blockDim.x) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    for (int r = b0; r < b1; ++r) {
      const float xv = __ldg(x + (size_t)r * k_in + kc);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < n_out) acc[j] = fmaf(__ldg(g + (size_t)r * n_out + j), xv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < n_out) atomicAdd(dw + (size_t)j * k_in + kc, acc[j]);
  }
}

}  // namespace lt_wgrad

using lt_wgrad::launch;
using lt_wgrad::wgrad_narrow_kernel;
using lt_wgrad::wgrad_narrow_scalar_kernel;

// dw[n_out, k_in] (+)= g[B, n_out]^T x[B, k_in];  dbias[n_out] (+)= column sums of g (dbias may be null).  zero_first != 0: dw and dbias
// are cleared on the stream first (memset nodes under capture); otherwise the caller has cleared them (PPO clears the whole flat
// gradient buffer once per mini-batch).
extern "C" int lt_wgrad_splitk(const float* grad_out, const float* act_in, float* dw, float* dbias, int B, int n_out, int k_in, int zero_first, void* stream) {
  if (!grad_out || !act_in || !dw || B <= 0 || n_out <= 0 || k_in <= 0) return LT_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool narrow = n_out <= 16;
  if (!narrow && ((n_out & 3) || (k_in & 3) || (((uintptr_t)grad_out | (uintptr_t)act_in | (uintptr_t)dw) & 15))) return LT_ERR_UNSUPPORTED;  // TMA: 16-byte rows
  if (zero_first) {
    if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)n_out * k_in, st) != cudaSuccess) return LT_ERR_CUDA;
    if (dbias && cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)n_out, st) != cudaSuccess) return LT_ERR_CUDA;
  }
  if (narrow) {
    const int blocks = 2 * lt::sm_count();
    const int rows = (B + blocks - 1) / blocks;
    const int grid = (B + rows - 1) / rows;
    if ((k_in & 3) == 0 && k_in <= 1024 && (((uintptr_t)act_in | (uintptr_t)dw) & 15) == 0)
      (n_out <= 4 ? wgrad_narrow_kernel<4> : wgrad_narrow_kernel<16>)<<<grid, 256, 0, st>>>(grad_out, act_in, dw, dbias, B, n_out, k_in, rows);
    else
      wgrad_narrow_scalar_kernel<<<grid, 256, 0, st>>>(grad_out, act_in, dw, dbias, B, n_out, k_in, rows);
    return lt::check_launch();
  }
  static const int knob_ctas = getenv("LT_WGRAD_CTAS") ? atoi(getenv("LT_WGRAD_CTAS")) : 0;
  static const int knob_wide = getenv("LT_WGRAD_WIDE") ? atoi(getenv("LT_WGRAD_WIDE")) : 1;
  static const int bulk = getenv("LT_WGRAD_BULK") ? atoi(getenv("LT_WGRAD_BULK")) : 1;
  const int sms = knob_ctas > 0 ? knob_ctas : lt::sm_count();
  if (!knob_wide) {
    if (k_in > 128) return launch<256, 1>(grad_out, act_in, dw, dbias, B, n_out, k_in, sms, bulk, st);
    if (k_in > 64) return launch<128, 1>(grad_out, act_in, dw, dbias, B, n_out, k_in, sms, bulk, st);
    return launch<64, 1>(grad_out, act_in, dw, dbias, B, n_out, k_in, sms, bulk, st);
  }
  // one CTA spans the full width of dW whenever it fits the 512 TMEM columns (k_in <= 512); wider layers tile the columns too
  if (k_in > 384) return launch<256, 2>(grad_out, act_in, dw, dbias, B, n_out, k_in, sms, bulk, st);
  if (k_in > 256) return launch<192, 2>(grad_out, act_in, dw, dbias, B, n_out, k_in, sms, bulk, st);
  if (k_in > 128) return launch<256, 1>(grad_out, act_in, dw, dbias, B, n_out, k_in, sms, bulk, st);
  if (k_in > 64) return launch<128, 1>(grad_out, act_in, dw, dbias, B, n_out, k_in, sms, bulk, st);
  return launch<64, 1>(grad_out, act_in, dw, dbias, B, n_out, k_in, sms, bulk, st);
}

// The same for TWO layers of identical shape (the actor's and the critic's layer i) in ONE launch: the CTAs are divided between the two
// problems, so the fixed cost of a launch (prologue, pipeline fill, the flush of the 128-row accumulator slabs) is paid once and every
// CTA accumulates twice as many batch rows before it adds its slab into dW.  Outputs must have been cleared by the caller.
extern "C" int lt_wgrad_splitk_pair(const float* grad_out0, const float* act_in0, float* dw0, float* dbias0, const float* grad_out1, const float* act_in1, float* dw1,
                                    float* dbias1, int B, int n_out, int k_in, void* stream) {
  if (!grad_out0 || !act_in0 || !dw0 || !grad_out1 || !act_in1 || !dw1 || B <= 0 || n_out <= 0 || k_in <= 0 || (dbias0 == nullptr) != (dbias1 == nullptr))
    return LT_ERR_INVALID_ARG;
  if (n_out <= 16 || (n_out & 3) || (k_in & 3) ||
      (((uintptr_t)grad_out0 | (uintptr_t)act_in0 | (uintptr_t)dw0 | (uintptr_t)grad_out1 | (uintptr_t)act_in1 | (uintptr_t)dw1) & 15))
    return LT_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = lt::sm_count();
  if (k_in > 384) return launch<256, 2>(grad_out0, act_in0, dw0, dbias0, B, n_out, k_in, sms, 1, st, grad_out1, act_in1, dw1, dbias1);
  if (k_in > 256) return launch<192, 2>(grad_out0, act_in0, dw0, dbias0, B, n_out, k_in, sms, 1, st, grad_out1, act_in1, dw1, dbias1);
  if (k_in > 128) return launch<256, 1>(grad_out0, act_in0, dw0, dbias0, B, n_out, k_in, sms, 1, st, grad_out1, act_in1, dw1, dbias1);
  if (k_in > 64) return launch<128, 1>(grad_out0, act_in0, dw0, dbias0, B, n_out, k_in, sms, 1, st, grad_out1, act_in1, dw1, dbias1);
  return launch<64, 1>(grad_out0, act_in0, dw0, dbias0, B, n_out, k_in, sms, 1, st, grad_out1, act_in1, dw1, dbias1);
}

#else  // !LT_HAVE_CUTLASS

extern "C" int lt_wgrad_splitk_pair(const float*, const float*, float*, float*, const float*, const float*, float*, float*, int, int, int, void*) {
  return LT_ERR_UNSUPPORTED;
}
extern "C" int lt_wgrad_splitk(const float*, const float*, float*, float*, int, int, int, int, void*) { return LT_ERR_UNSUPPORTED; }

#endif
