// K19: the three hidden layers of the actor / critic MLP (k0 -> 512 -> 256 -> 128, bias + ELU after each) of BOTH networks as ONE
// persistent hand-written tcgen05 kernel.  Replaces, per rollout step (reference loco_rl/loco_rl/modules/actor_critic.py:105-131
// `act` / `evaluate`, called from algorithms/ppo.py:129-141) and per mini-batch forward (algorithms/ppo.py:264-281 ->
// actor_critic.py:33-56), six K12 launches (one GEMM per layer and network) whose activations travel through HBM / L2 between them.
//
// One CTA owns a 128-row slab of one network through all three layers; the activations never leave the SM:
//   * layer 1: the slab of x ([128 x k0], <= 176 KB) is loaded ONCE into shared memory (TMA, K-major SW128 boxes of [128 x 32]) and
//     multiplied against W1 in four column chunks of 128: D1 chunk c (TMEM, 128 columns, double-buffered) = x . W1[c*128:(c+1)*128]^T,
//     both operands from shared memory (tcgen05.mma kind::tf32, M = 128, N = 128, K = 8);
//   * the epilogue warps read a finished chunk (tcgen05.ld), add the bias, apply ELU and write it back IN PLACE (tcgen05.st): the fp32
//     accumulator layout (lane = row, column = n) is exactly the layout of a K-major 32-bit A operand in TMEM, so the chunk IS the A
//     operand of layer 2 for k = c*128 .. c*128+127 -- no shared memory, no swizzle, no copy;
//   * layer 2: D2 (256 TMEM columns, two N halves) += A2 chunk (TMEM) . W2[:, chunk]^T as soon as a chunk is ready, interleaved with the
//     layer-1 MMAs of the next chunk, so the tensor pipe runs while the epilogue warps work;
//   * layer 3: D3 (128 columns, over the idle chunk buffer) = ELU(D2 + b2) (in place, TMEM) . W3^T; its epilogue writes h3.
// TMEM: 2 x 128 (D1 chunks / A2 / D3) + 256 (D2 / A3) = 512 columns.  Shared memory: x slab 11 x 16 KB + a 48 KB ring of [128 x 32]
// weight tiles (84 tiles per slab in the fixed order the MMA warp consumes them) -- the weights (1.37 MB per network) are
// L2 resident and are the only operand that streams.  Per-SM fill is 1.55 MB per slab against 2.1 MB (x re-read per column tile) for
// three separate GEMMs, and the per-layer launch tails, the activation round trips (write h, re-read h) and the non-overlapped
// epilogues of the per-layer kernels are gone.  Training mode additionally stores h1 / h2 (the backward pass needs them) straight
// from the epilogue registers.
// Warp roles (20 warps): warp 0 = weight-tile producer, warp 1 = MMA issuer, warp 2 = x producer, warp 3 idle, warps 4-19 = epilogue
// (TMEM lane quarter = warp % 4, four warps per quarter split a chunk's columns).  Producer and MMA roles run warp-uniform with only the
// TMA / MMA / commit instructions under elect_one (a single diverged lane pays a register -> uniform-register move per MMA operand).
// Pair mode (template parameter, the default): the two CTAs of a cluster carry two adjacent slabs of one network; every weight tile is
// split between their shared memories (64 of its 128 rows each: six ring slots instead of three, half the weight traffic per SM), the
// leader CTA issues cta_group::2 MMAs (M = 256) for both, the TMA bytes of both halves are counted on the leader's barriers, commits are
// multicast to both CTAs and the peer's epilogue warps arrive remotely on the leader's barriers.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "lt_common.cuh"
#include "locotouch_b200.h"

#if LT_HAVE_CUTLASS

#include <cute/tensor.hpp>
#include <cute/arch/copy_sm90_desc.hpp>
#include <cute/arch/copy_sm90_tma.hpp>
#include <cute/atom/copy_traits_sm90_tma.hpp>
#include <cute/atom/mma_traits_sm100.hpp>

namespace lt_mlp3 {

using namespace cute;

constexpr int kRows = 128;           // slab rows = UMMA M
constexpr int kH1 = 512, kH2 = 256, kH3 = 128;
constexpr int kBK = 32;              // fp32 columns of one SW128 row (128 bytes)
constexpr int kTile = kRows * kBK * 4;  // 16 KB: one [128 x 32] fp32 box
constexpr int kMaxKB0 = 11;          // k0 <= 352
constexpr int kWK = 32;              // fp32 columns of one weight tile: [128 x 32], 128-byte rows (SW128), 16 KB
constexpr int kWTile = kRows * kWK * 4;
constexpr int kSlots = 3;            // weight-tile ring (one CTA per slab)
constexpr int kSlotsPair = 6;        // pair mode: a CTA holds half of every weight tile (64 rows, 8 KB), so the same 48 KB are six slots
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;  // warp 0 weight producer, 1 MMA issuer, 2 x producer, 3 idle, 4-19 epilogue
constexpr uint32_t kColD1 = 0;       // two chunk buffers: columns 0-127, 128-255
constexpr uint32_t kColD2 = 256;     // 256 columns
constexpr uint32_t kColD3 = 128;     // over chunk buffer 1

struct Bars {
  uint64_t x_full[kMaxKB0];
  uint64_t x_empty;
  uint64_t ring_full[kSlotsPair];
  uint64_t ring_empty[kSlotsPair];
  uint64_t d1_full[4];
  uint64_t a2_ready[4];
  uint64_t d2_full, a3_ready, d3_full, d3_empty;
  uint32_t tmem_base;
};
constexpr int kSmemBytes = kMaxKB0 * kTile + kSlots * kWTile + 1024 + (int)sizeof(Bars);
static_assert(kSmemBytes <= 232448, "shared memory budget");

struct NetArgs {
  const float* b1;
  const float* b2;
  const float* b3;
  float* h1;   // [B, 512] or null (rollout: only h3 is needed)
  float* h2;   // [B, 256] or null
  float* h3;   // [B, 128]
};
struct alignas(64) Params {
  CUtensorMap x[2], w1[2], w2[2], w3[2];
  NetArgs net[2];
  int n_nets, B, nkb0, slabs, pairs_per_net;
  unsigned long long* dbg;  // LT_MLP3_DBG: per-CTA wait-cycle attribution (debug builds of the launcher only)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint64_t hint) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor of a K-major [rows x 32 fp32] SW128 box: 8-row groups of 1024 bytes (SBO), version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// the same for a [rows x 16 fp32] SW64 box: 8-row groups of 512 bytes, SWIZZLE_64B
__device__ __forceinline__ uint64_t smem_desc_w(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(kIdesc),
               "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b),
               "r"(kIdesc), "r"(acc)
               : "memory");
}

// The MMA issuer is ONE thread: its own instruction stream (~5 cycles per dependent instruction) must stay below the 82 cycles a
// 128 x 128 x 8 MMA takes, so descriptors are kept as 32-bit low words (start address >> 4) that are advanced by adding constants;
// the high words are compile-time constants.
constexpr uint32_t kDescHiSW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t kDescHiSW64 = (512u >> 4) | (1u << 14) | (4u << 29);
constexpr uint32_t kDescHiW = kWK == 32 ? kDescHiSW128 : kDescHiSW64;  // weight tiles
__device__ __forceinline__ void mma_ss_lo(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}" ::"r"(d),
      "r"(a_lo), "r"(b_lo), "r"(kIdesc), "r"(acc), "r"(kDescHiSW128), "r"(kDescHiW)
      : "memory");
}
__device__ __forceinline__ void mma_ts_lo(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "r"(b_lo), "r"(kIdesc), "r"(acc), "r"(kDescHiW)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar_addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}

// ---- pair mode (cta_group::2): two CTAs of a cluster carry two adjacent slabs of one network; every weight tile is split between
// their shared memories (64 of its 128 rows each), the leader CTA issues M = 256 MMAs for both.
constexpr uint32_t kIdescPair = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared-memory address: the same offset in the leader CTA
__device__ __forceinline__ void mma2_ss_lo(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %3, p;\n\t}" ::"r"(d),
      "r"(a_lo), "r"(b_lo), "r"(kIdescPair), "r"(acc), "r"(kDescHiSW128), "r"(kDescHiW)
      : "memory");
}
__device__ __forceinline__ void mma2_ts_lo(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], db, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "r"(b_lo), "r"(kIdescPair), "r"(acc), "r"(kDescHiW)
      : "memory");
}
__device__ __forceinline__ void umma_commit2_a(uint32_t bar_addr) {  // arrives on the barrier at this offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_a(uint32_t bar_addr, uint32_t parity) {  // arrivals may come from the peer CTA
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar_addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar_addr, bool remote) {  // arrive on the LEADER's barrier at this offset
  if (remote) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\tmbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar_addr) : "memory");
  } else {
    asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <bool kPair>
__device__ __forceinline__ void tma_load_t(const CUtensorMap* map, uint32_t bar_addr, uint32_t dst_addr, int c0, int c1, uint64_t hint) {
  if constexpr (kPair) {  // both CTAs load; the bytes are counted on the leader's barrier
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst_addr),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr & kPeerMask), "r"(c0), "r"(c1), "l"(hint)
                 : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst_addr),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1), "l"(hint)
                 : "memory");
  }
}

#define LT_TMEM_LD32(r, taddr)                                                                                                                         \
  asm volatile(                                                                                                                                        \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                                                        \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                                                        \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                        \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),        \
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),          \
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),          \
        "=r"(r[31])                                                                                                                                    \
      : "r"(taddr))
#define LT_TMEM_ST32(r, taddr)                                                                                                                         \
  asm volatile(                                                                                                                                        \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "                                                                                                 \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                                                        \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};" ::"r"(r[0]),                                                  \
      "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),         \
      "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),            \
      "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)                                     \
      : "memory")

// torch.nn.functional.elu (alpha = 1) on a TF32-GEMM output: x > 0 ? x : exp(x) - 1
__device__ __forceinline__ float elu1(float v) {
  float t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(v * 1.4426950408889634f));  // one MUFU, no denormal-range fix-up code
  return v > 0.0f ? v : t - 1.0f;
}

// One 32-column block of an accumulator: TMEM -> registers, + bias, ELU (in r), optionally back to TMEM in place.
__device__ __forceinline__ void epi_compute(uint32_t (&r)[32], uint32_t taddr, const float* __restrict__ bias, bool write_back, long long* prof) {
  const long long t0 = prof ? clock64() : 0;
  LT_TMEM_LD32(r, taddr);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const long long t1 = prof ? clock64() : 0;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + j));
    r[j] = __float_as_uint(elu1(__uint_as_float(r[j]) + b.x));
    r[j + 1] = __float_as_uint(elu1(__uint_as_float(r[j + 1]) + b.y));
    r[j + 2] = __float_as_uint(elu1(__uint_as_float(r[j + 2]) + b.z));
    r[j + 3] = __float_as_uint(elu1(__uint_as_float(r[j + 3]) + b.w));
  }
  if (prof) {
    asm volatile("" ::"r"(r[0]), "r"(r[7]), "r"(r[15]), "r"(r[23]), "r"(r[31]) : "memory");
    prof[1] += t1 - t0;
    prof[2] += clock64() - t1;
  }
  if (write_back) LT_TMEM_ST32(r, taddr);
}

// the block's row to global memory: 32-byte stores, every lane writes whole sectors of its own row (rows are 128-byte aligned)
__device__ __forceinline__ void epi_store(const uint32_t (&r)[32], float* __restrict__ out_row, long long* prof) {
  const long long t0 = prof ? clock64() : 0;
#pragma unroll
  for (int j = 0; j < 32; j += 8)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out_row + j), "r"(r[j]), "r"(r[j + 1]), "r"(r[j + 2]), "r"(r[j + 3]), "r"(r[j + 4]),
                 "r"(r[j + 5]), "r"(r[j + 6]), "r"(r[j + 7])
                 : "memory");
  if (prof) prof[3] += clock64() - t0;
}

// epilogue of `nblk` 32-column blocks of one accumulator for this thread's row.  The hand-over to the MMA warp (the accumulator is
// written back / drained) comes BEFORE the global stores of the last block: a full store queue then holds up this warp, not the tensor pipe.
template <bool kPair>
__device__ __forceinline__ void epi_job(uint32_t wait_bar, uint32_t par, uint32_t taddr, const float* bias, float* out, int nblk, bool write_back, uint32_t arrive_bar,
                                        long long* wait_cycles, bool remote) {
  {
    const long long t0 = wait_cycles ? clock64() : 0;
    mbar_wait_a(wait_bar, par);
    if (wait_cycles) *wait_cycles += clock64() - t0;
  }
  fence_after();
  uint32_t r[32];
#pragma unroll 1
  for (int blk = 0; blk + 1 < nblk; ++blk) {
    epi_compute(r, taddr + blk * 32, bias + blk * 32, write_back, wait_cycles);
    if (out) epi_store(r, out + blk * 32, wait_cycles);
  }
  const int last = nblk - 1;
  epi_compute(r, taddr + last * 32, bias + last * 32, write_back, wait_cycles);
  if (write_back) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    if constexpr (kPair)
      mbar_arrive_leader(arrive_bar, remote);
    else
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(arrive_bar) : "memory");
  }
  if (out) epi_store(r, out + last * 32, wait_cycles);
}

template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1) mlp3_forward_kernel(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = base;                               // [kb][128 x 32] SW128 boxes of the x slab
  uint8_t* ring = base + kMaxKB0 * kTile;           // SW128 weight tiles: [128 x 32], or this CTA's [64 x 32] half in pair mode
  Bars& bars = *reinterpret_cast<Bars*>(ring + kSlots * kWTile);
  constexpr uint32_t NS = kPair ? kSlotsPair : kSlots;            // ring slots
  constexpr uint32_t kSlotBytes = kPair ? kWTile / 2 : kWTile;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (kPair) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = rank == 0;
  // tasks of this CTA: slab `slab_of(task)` of network `net_of(task)`; in pair mode a task is two adjacent slabs (one per CTA)
  const int first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tasks = kPair ? P.n_nets * P.pairs_per_net : P.n_nets * P.slabs;
  auto net_of = [&](int task) { return kPair ? task / P.pairs_per_net : task / P.slabs; };
  auto slab_of = [&](int task) { return kPair ? 2 * (task % P.pairs_per_net) + (int)rank : task % P.slabs; };
  const int nkb0 = P.nkb0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxKB0; ++i) mbar_init(&bars.x_full[i], 1);
    mbar_init(&bars.x_empty, 1);
    for (int i = 0; i < (int)NS; ++i) {
      mbar_init(&bars.ring_full[i], 1);
      mbar_init(&bars.ring_empty[i], 1);
    }
    constexpr int kEpiArrivals = kPair ? 2 * kEpiWarps : kEpiWarps;  // pair mode: the leader's MMA warp waits for the epilogues of both CTAs
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars.d1_full[i], 1);
      mbar_init(&bars.a2_ready[i], kEpiArrivals);
    }
    mbar_init(&bars.d2_full, 1);
    mbar_init(&bars.a3_ready, kEpiArrivals);
    mbar_init(&bars.d3_full, 1);
    mbar_init(&bars.d3_empty, kEpiArrivals);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars.tmem_base)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars.tmem_base)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  fence_before();
  if constexpr (kPair) cluster_sync_all(); else __syncthreads();
  fence_after();
  const uint32_t tmem = bars.tmem_base;
  const uint32_t full0 = smem_u32(&bars.ring_full[0]), empty0 = smem_u32(&bars.ring_empty[0]);

  // Producer and MMA roles are executed by their WHOLE warp with warp-uniform control flow; only the TMA / MMA / commit instructions
  // themselves sit under elect_one: descriptor arithmetic then lives in the uniform datapath (a single diverged lane would pay a
  // register -> uniform-register move per operand of every tcgen05.mma).
  if (warp == 0) {
    // ------------------------------------------------------------------ weight-tile producer: the tiles in the order of the MMA warp
    const uint64_t keep = (uint64_t)TMA::CacheHintSm90::EVICT_LAST;
    const uint32_t ring_addr = smem_u32(ring);
    uint32_t slot = 0, phase = 0, wrapped = 0;
    auto tiles = [&](const CUtensorMap* map, int col0, int row, int n) {
#pragma unroll 1
      for (int t = 0; t < n; ++t) {
        if (wrapped) mbar_wait_a(empty0 + slot * 8, phase ^ 1);
        if (cute::elect_one_sync()) {
          const uint32_t bar = full0 + slot * 8;
          // the whole tile (both halves in pair mode) is counted on the leader's barrier
          if (leader) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kWTile) : "memory");
          tma_load_t<kPair>(map, bar, ring_addr + slot * kSlotBytes, col0 + t * kWK, row + (kPair ? (int)rank * 64 : 0), keep);
        }
        __syncwarp();
        if (++slot == NS) {
          slot = 0;
          phase ^= 1;
          wrapped = 1;
        }
      }
    };
#pragma unroll 1
    for (int task = first; task < tasks; task += stride) {
      const int net = net_of(task);
      if (task == first) tiles(&P.w1[net], 0, 0, nkb0);
      tiles(&P.w1[net], 0, 128, nkb0);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {  // layer-2 part j (two N halves), then layer-1 chunk j + 2: the order of the MMA warp
        tiles(&P.w2[net], j * 128, 0, 128 / kWK);
        tiles(&P.w2[net], j * 128, 128, 128 / kWK);
        if (j < 2) tiles(&P.w1[net], 0, (j + 2) * 128, nkb0);
      }
      if (task + stride < tasks) tiles(&P.w1[net_of(task + stride)], 0, 0, nkb0);  // chunk 0 of the next slab runs ahead of layer 3
      tiles(&P.w3[net], 0, 0, 256 / kWK);
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ x producer: the whole slab as soon as the buffer is free
    const uint64_t once = (uint64_t)TMA::CacheHintSm90::EVICT_FIRST;
    int it = 0;
#pragma unroll 1
    for (int task = first; task < tasks; task += stride, ++it) {
      const int net = net_of(task), row0 = slab_of(task) * kRows;  // a padding slab (row0 >= B) reads zeros
      if (it > 0) mbar_wait(&bars.x_empty, (it - 1) & 1);
      if (cute::elect_one_sync()) {
#pragma unroll 1
        for (int kb = 0; kb < nkb0; ++kb) {
          if (leader) mbar_expect_tx(&bars.x_full[kb], kPair ? 2 * kTile : kTile);  // pair mode: the boxes of both CTAs, counted on the leader
          tma_load_t<kPair>(&P.x[net], smem_u32(&bars.x_full[kb]), smem_u32(xs + kb * kTile), kb * kBK, row0, once);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------------ MMA issuer (pair mode: the leader CTA issues for both)
    const uint32_t xs_lo = smem_u32(xs) >> 4, ring_lo = smem_u32(ring) >> 4;
    const uint32_t xfull0 = smem_u32(&bars.x_full[0]);
    const uint32_t d1f0 = smem_u32(&bars.d1_full[0]), a2r0 = smem_u32(&bars.a2_ready[0]);
    const uint32_t b_x_empty = smem_u32(&bars.x_empty), b_d3_empty = smem_u32(&bars.d3_empty), b_d2_full = smem_u32(&bars.d2_full),
                   b_a3_ready = smem_u32(&bars.a3_ready), b_d3_full = smem_u32(&bars.d3_full);
    const bool dbg = P.dbg != nullptr;
    long long w_ring = 0, w_ring_ts = 0, w_x = 0, w_epi = 0;
    const long long t_begin = clock64();
    uint32_t slot = 0, phase = 0;
    auto wait_t = [&](uint32_t bar, uint32_t parity, long long& acc) {
      const long long t0 = dbg ? clock64() : 0;
      if constexpr (kPair) mbar_wait_cluster_a(bar, parity); else mbar_wait_a(bar, parity);
      if (dbg) acc += clock64() - t0;
    };
    auto commit = [&](uint32_t bar) {
      if constexpr (kPair) umma_commit2_a(bar); else umma_commit_a(bar);
    };
    // `ntiles` weight tiles of one accumulation chain; a = low descriptor word of the first x box (SS) or TMEM address of the first A
    // column (TS); commits c1 / c2 (0: none) follow the last tile
    auto run = [&](bool is_ts, uint32_t ntiles, uint32_t a, uint32_t a_step, uint32_t d, uint32_t acc, uint32_t xbar, uint32_t xpar, uint32_t c1, uint32_t c2) {
#pragma unroll 1
      for (uint32_t t = 0; t < ntiles; ++t) {
        if (xbar) wait_t(xbar + t * 8, xpar, w_x);
        wait_t(full0 + slot * 8, phase, is_ts ? w_ring_ts : w_ring);
        fence_after();
        const uint32_t b = ring_lo + slot * (kSlotBytes >> 4);
        if (cute::elect_one_sync()) {
          if constexpr (kPair) {
            if (is_ts) {
              mma2_ts_lo(d, a, b, acc);
              mma2_ts_lo(d, a + 8, b + 2, 1);
              mma2_ts_lo(d, a + 16, b + 4, 1);
              mma2_ts_lo(d, a + 24, b + 6, 1);
            } else {
              mma2_ss_lo(d, a, b, acc);
              mma2_ss_lo(d, a + 2, b + 2, 1);
              mma2_ss_lo(d, a + 4, b + 4, 1);
              mma2_ss_lo(d, a + 6, b + 6, 1);
            }
          } else if (is_ts) {
            mma_ts_lo(d, a, b, acc);
            mma_ts_lo(d, a + 8, b + 2, 1);
            mma_ts_lo(d, a + 16, b + 4, 1);
            mma_ts_lo(d, a + 24, b + 6, 1);
          } else {
            mma_ss_lo(d, a, b, acc);
            mma_ss_lo(d, a + 2, b + 2, 1);
            mma_ss_lo(d, a + 4, b + 4, 1);
            mma_ss_lo(d, a + 6, b + 6, 1);
          }
          commit(empty0 + slot * 8);
          if (t + 1 == ntiles) {
            if (c1) commit(c1);
            if (c2) commit(c2);
          }
        }
        __syncwarp();
        if (++slot == NS) {
          slot = 0;
          phase ^= 1;
        }
        a += a_step;
        acc = 1;
      }
    };
    // layer 1, column chunk c of slab iteration `sit`: D1[c & 1] = x . W1[c*128 .. c*128+127]^T
    auto layer1_chunk = [&](int c, int sit) {
      if (c == 1 && sit > 0) {  // chunk buffer 1 held D3 of the previous slab
        wait_t(b_d3_empty, (sit - 1) & 1, w_epi);
        fence_after();
      }
      run(false, nkb0, xs_lo, kTile >> 4, tmem + kColD1 + (c & 1) * 128, 0, c == 0 ? xfull0 : 0, sit & 1, d1f0 + c * 8, c == 3 ? b_x_empty : 0);
    };
    auto layer2_part = [&](int j, uint32_t par) {
      wait_t(a2r0 + j * 8, par, w_epi);
      fence_after();
      const uint32_t a = tmem + kColD1 + (j & 1) * 128;
      run(true, 128 / kWK, a, 32, tmem + kColD2, j != 0, 0, 0, 0, 0);
      run(true, 128 / kWK, a, 32, tmem + kColD2 + 128, j != 0, 0, 0, j == 3 ? b_d2_full : 0, 0);
    };
    int it = 0;
#pragma unroll 1
    for (int task = first; task < tasks; task += stride, ++it) {
      const uint32_t par = it & 1;
      if (it == 0) layer1_chunk(0, 0);
      layer1_chunk(1, it);
      layer2_part(0, par);
      layer1_chunk(2, it);
      layer2_part(1, par);
      layer1_chunk(3, it);
      layer2_part(2, par);
      layer2_part(3, par);
      if (task + stride < tasks) layer1_chunk(0, it + 1);  // keeps the tensor pipe busy while the epilogue turns D2 into A3
      wait_t(b_a3_ready, par, w_epi);
      fence_after();
      run(true, 256 / kWK, tmem + kColD2, 32, tmem + kColD3, 0, 0, 0, b_d3_full, 0);
    }
    if (dbg && lane == 0) {
      unsigned long long* d = P.dbg + 8 * blockIdx.x;
      d[0] = (unsigned long long)(clock64() - t_begin);
      d[1] = (unsigned long long)(w_ring + w_ring_ts);
      P.dbg[8 * (blockIdx.x + 1) + 0] = (unsigned long long)w_ring_ts;  // pair mode only: the peer CTA's row is otherwise unused
      d[2] = (unsigned long long)w_x;
      d[3] = (unsigned long long)w_epi;
      d[4] = (unsigned long long)it;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps
    const int e = warp - 4;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int part = e >> 2;           // which quarter of a chunk's columns (0..3)
    const int r_local = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t d1f0 = smem_u32(&bars.d1_full[0]), a2r0 = smem_u32(&bars.a2_ready[0]);
    long long w_prof[4] = {0, 0, 0, 0};  // wait for the MMA warp, TMEM load, bias + ELU, TMEM store + global stores
    long long& w_mma = w_prof[0];
    long long* wp = P.dbg ? w_prof : nullptr;
    const long long t_begin = clock64();
    int it = 0;
    const bool remote = kPair && !leader;
#pragma unroll 1
    for (int task = first; task < tasks; task += stride, ++it) {
      const uint32_t par = it & 1;
      const int net = net_of(task), row = slab_of(task) * kRows + r_local;
      const NetArgs& na = P.net[net];
      const bool live = row < P.B;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col = c * 128 + part * 32;
        epi_job<kPair>(d1f0 + c * 8, par, lane_addr + kColD1 + (c & 1) * 128 + part * 32, na.b1 + col,
                       (na.h1 != nullptr && live) ? na.h1 + (size_t)row * kH1 + col : nullptr, 1, true, a2r0 + c * 8, wp, remote);
      }
      epi_job<kPair>(smem_u32(&bars.d2_full), par, lane_addr + kColD2 + part * 64, na.b2 + part * 64,
                     (na.h2 != nullptr && live) ? na.h2 + (size_t)row * kH2 + part * 64 : nullptr, 2, true, smem_u32(&bars.a3_ready), wp, remote);
      epi_job<kPair>(smem_u32(&bars.d3_full), par, lane_addr + kColD3 + part * 32, na.b3 + part * 32, live ? na.h3 + (size_t)row * kH3 + part * 32 : nullptr, 1, false,
                     smem_u32(&bars.d3_empty), wp, remote);
    }
    if (P.dbg && warp == 4 && lane == 0) {
      P.dbg[8 * blockIdx.x + 5] = (unsigned long long)(clock64() - t_begin);
      P.dbg[8 * blockIdx.x + 6] = (unsigned long long)w_mma;
      P.dbg[8 * blockIdx.x + 7] = (unsigned long long)w_prof[1] | ((unsigned long long)w_prof[2] << 20) | ((unsigned long long)w_prof[3] << 40);
    }
  }
  fence_before();
  if constexpr (kPair) {  // the peer's shared memory, TMEM and barriers are in use until the leader's last MMA has completed
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  } else {
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}


// Measurement aid (LT_MLP3_RATE): issue rate of the MMAs this kernel is made of, on garbage operands: cycles per instruction for
// SS / TS and N = 128 / 256, one CTA per SM.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(unsigned long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2, 1 << 20);
    mbar_init(&bar3, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 3 * kTile / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(base), b = smem_u32(base) + kTile;
    constexpr uint32_t idesc256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t phase = 0;
    for (int mode = 0; mode < 8; ++mode) {
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint32_t k = (i & 3) * 32;
        if (mode == 0 || (mode >= 4 && mode <= 6)) mma_ss(tmem, smem_desc(a + k), smem_desc(b + k), 1);
        if (mode == 7) mma_ss(tmem, smem_desc_w(a + (k & 32)), smem_desc_w(b + (k & 32)), 1);
        if ((mode == 4 || mode == 6) && (i & 3) == 3) umma_commit(&bar2);
        if (mode == 5 && (i & 1) == 1) umma_commit(&bar2);
        if (mode == 6 && (i & 3) == 3) {
          mbar_wait(&bar3, 1);
          fence_after();
        }
        if (mode == 1) mma_ts(tmem + 256, tmem + (i & 15) * 8, smem_desc(b + k), 1);
        if (mode == 2)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(smem_desc(a + k)),
                       "l"(smem_desc(b + k)), "r"(idesc256), "r"(1)
                       : "memory");
        if (mode == 3)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem + 256),
                       "r"(tmem + (i & 15) * 8), "l"(smem_desc(b + k)), "r"(idesc256), "r"(1)
                       : "memory");
      }
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
      if (blockIdx.x == 0) out[mode] = (unsigned long long)(clock64() - t0);
    }
  }
  fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// K-major [rows, k] fp32 tensor -> tensor map with [128 x 32] SW128 boxes (rows / columns beyond the tensor read as zero)
template <int BK, int ROWS = kRows>
static bool make_map(CUtensorMap* out, const float* ptr, int rows, int k) {
  using TF = cute::tfloat32_t;
  using Atom = cute::conditional_t<BK == 32, UMMA::Layout_K_SW128_Atom<TF>, UMMA::Layout_K_SW64_Atom<TF>>;
  Tensor g = make_tensor(make_gmem_ptr(reinterpret_cast<TF const*>(ptr)), make_layout(make_shape(rows, k), make_stride(k, Int<1>{})));
  auto sl = tile_to_shape(Atom{}, make_shape(Int<ROWS>{}, Int<BK>{}));
  auto atom = make_tma_atom(SM90_TMA_LOAD{}, g, sl, make_shape(Int<ROWS>{}, Int<BK>{}));
  static_assert(sizeof(*atom.get_tma_descriptor()) == sizeof(CUtensorMap), "descriptor size");
  memcpy(out, atom.get_tma_descriptor(), sizeof(CUtensorMap));
  return true;
}

}  // namespace lt_mlp3

extern "C" int lt_mlp3_forward(const LtMlp3Net* nets, int n_nets, int B, void* stream) {
  using namespace lt_mlp3;
  if (!nets || n_nets < 1 || n_nets > 2 || B <= 0) return LT_ERR_INVALID_ARG;
  const int k0 = nets[0].k0;
  if (k0 <= 0 || (k0 & 3) || k0 > kMaxKB0 * kBK) return LT_ERR_UNSUPPORTED;
  static const bool pair = !(getenv("LT_MLP3_PAIR") && atoi(getenv("LT_MLP3_PAIR")) == 0);  // cta_group::2 pairs (default) / one CTA per slab
  Params P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < n_nets; ++i) {
    const LtMlp3Net& n = nets[i];
    if (!n.x || !n.w1 || !n.b1 || !n.w2 || !n.b2 || !n.w3 || !n.b3 || !n.h3) return LT_ERR_INVALID_ARG;
    if (n.k0 != k0) return LT_ERR_UNSUPPORTED;
    const uintptr_t al = (uintptr_t)n.x | (uintptr_t)n.w1 | (uintptr_t)n.w2 | (uintptr_t)n.w3 | (uintptr_t)n.b1 | (uintptr_t)n.b2 | (uintptr_t)n.b3 |
                         (uintptr_t)n.h1 | (uintptr_t)n.h2 | (uintptr_t)n.h3;
    if ((al & 15) || (((uintptr_t)n.h1 | (uintptr_t)n.h2 | (uintptr_t)n.h3) & 31)) return LT_ERR_UNSUPPORTED;  // TMA: 16 bytes; 32-byte stores of the rows
    make_map<kBK>(&P.x[i], n.x, B, k0);
    if (pair) {  // a CTA loads its 64-row half of every weight tile
      make_map<kWK, 64>(&P.w1[i], n.w1, kH1, k0);
      make_map<kWK, 64>(&P.w2[i], n.w2, kH2, kH1);
      make_map<kWK, 64>(&P.w3[i], n.w3, kH3, kH2);
    } else {
      make_map<kWK>(&P.w1[i], n.w1, kH1, k0);
      make_map<kWK>(&P.w2[i], n.w2, kH2, kH1);
      make_map<kWK>(&P.w3[i], n.w3, kH3, kH2);
    }
    P.net[i] = NetArgs{n.b1, n.b2, n.b3, n.h1, n.h2, n.h3};
  }
  P.n_nets = n_nets;
  P.B = B;
  P.nkb0 = (k0 + kBK - 1) / kBK;
  P.slabs = (B + kRows - 1) / kRows;
  P.pairs_per_net = (P.slabs + 1) / 2;  // an odd slab count is padded with an all-zero slab that stores nothing
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(mlp3_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(mlp3_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  });
  if (attr_err != cudaSuccess) return LT_ERR_CUDA;
  static const int knob_ctas = getenv("LT_MLP3_CTAS") ? atoi(getenv("LT_MLP3_CTAS")) : 0;
  const int tasks = pair ? n_nets * P.pairs_per_net : n_nets * P.slabs;
  int grid = knob_ctas > 0 ? knob_ctas : lt::sm_count();
  if (pair) grid /= 2;                 // clusters
  if (grid > tasks) grid = tasks;
  if (grid < 1) grid = 1;
  if (pair) grid *= 2;
  auto launch = [&]() -> int {
    if (!pair) {
      mlp3_forward_kernel<false><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
      return lt::check_launch();
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, mlp3_forward_kernel<true>, P) != cudaSuccess) {
      lt::set_last_cuda_error(cudaGetLastError());
      return LT_ERR_CUDA;
    }
    return lt::check_launch();
  };
  if (getenv("LT_MLP3_RATE")) {
    unsigned long long* d = nullptr;
    unsigned long long h[8];
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kTile);
    for (int ctas : {1, 148}) {
      mma_rate_kernel<<<ctas, 128, 4 * kTile>>>(d, 2048);
      cudaDeviceSynchronize();
      cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[mlp3 rate] %d CTAs: cycles per MMA (K = 8, M = 128): SS N128 %.1f  TS N128 %.1f  SS N256 %.1f  TS N256 %.1f | SS N128 + commit/4 %.1f  commit/2 %.1f  "
              "commit+wait+fence/4 %.1f  SW64 operands %.1f\n", ctas, h[0] / 2048.0, h[1] / 2048.0, h[2] / 2048.0, h[3] / 2048.0, h[4] / 2048.0, h[5] / 2048.0, h[6] / 2048.0,
              h[7] / 2048.0);
    }
    cudaFree(d);
  }
  static const bool dbg = getenv("LT_MLP3_DBG") != nullptr;
  if (dbg) {  // wait-cycle attribution of the MMA thread, printed per launch (synchronises: measurement aid only)
    static unsigned long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 8 * sizeof(unsigned long long) * 1024);
    cudaMemset(dbuf, 0, 8 * sizeof(unsigned long long) * 1024);
    P.dbg = dbuf;
    const int rc = launch();
    cudaDeviceSynchronize();
    static unsigned long long host[8 * 1024];
    cudaMemcpy(host, dbuf, sizeof(unsigned long long) * 8 * grid, cudaMemcpyDeviceToHost);
    double tot = 0, ring = 0, ring_ts = 0, xw = 0, epi = 0, slabs = 0, etot = 0, ewait = 0, eld = 0, emath = 0, est = 0;
    int issuers = 0;
    for (int i = 0; i < grid; ++i) {
      if (pair && (i & 1)) { ring_ts += host[8 * i]; host[8 * i] = 0; }
      if (host[8 * i]) ++issuers;  // pair mode: only the leader CTAs issue MMAs
      tot += host[8 * i]; ring += host[8 * i + 1]; xw += host[8 * i + 2]; epi += host[8 * i + 3]; slabs += host[8 * i + 4];
      etot += host[8 * i + 5]; ewait += host[8 * i + 6];
      eld += host[8 * i + 7] & 0xFFFFF; emath += (host[8 * i + 7] >> 20) & 0xFFFFF; est += (host[8 * i + 7] >> 40) & 0xFFFFF;
    }
    if (issuers < 1) issuers = 1;
    fprintf(stderr, "[mlp3 dbg] B=%d grid=%d pair=%d tasks/issuer=%.2f  MMA warp cycles per issuing CTA: total %.0f  wait ring %.0f (layers 2-3: %.0f)  wait x %.0f  wait epilogue %.0f | "
            "epilogue warp (per CTA): total %.0f  wait MMA %.0f  TMEM ld %.0f  bias+ELU %.0f  TMEM st + global st %.0f\n", B, grid, (int)pair, slabs / issuers, tot / issuers,
            ring / issuers, ring_ts / issuers, xw / issuers, epi / issuers, etot / grid, ewait / grid, eld / grid, emath / grid, est / grid);
    return rc;
  }
  return launch();
}

#else  // !LT_HAVE_CUTLASS

extern "C" int lt_mlp3_forward(const LtMlp3Net*, int, int, void*) { return LT_ERR_UNSUPPORTED; }

#endif
