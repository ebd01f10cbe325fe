// Shared device-side primitives for the sm_100a kernels of the ViLBERT hot path:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), UMMA
// descriptors, Philox dropout and small vector helpers.  Everything is inline PTX;
// no CUTLASS / CuTe types are used.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vilbert_b200.h"  // vb_status codes returned across the C ABI

namespace vb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded by WALL CLOCK: a try_wait may itself sleep for a hardware-defined time, so a spin COUNT bounds nothing
  // (a lost TMA completion once held a GPU box for 15 minutes).  2 s of %globaltimer, checked every 1024 polls.
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: c0 = coordinate along the contiguous dimension, c1 = row coordinate
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// D[global] += smem tile (element type comes from the tensor map: fp32 here); split-K / accumulating outputs
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-collective forms: the WHOLE (converged) warp executes the call with warp-uniform operands and one elected lane
// issues.  Unlike `if (lane == 0) umma_bf16(...)`, the call site is not divergent, so ptxas keeps the operands in uniform
// registers and emits a straight UTCHMMA instead of an ELECT / R2UR / BRA.U.ANY loop around every instruction.
__device__ __forceinline__ void umma_bf16_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a 2-wide cluster on the two SMs of a TPC run ONE UMMA of M = 256; each CTA holds
// its own 128 rows of A and HALF of the B tile in shared memory and its own 128 accumulator rows in TMEM.  The even CTA
// ("leader") issues the MMAs; TMA loads of both CTAs signal the leader's barrier, commits are multicast to both.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// execution-only cluster barrier (no memory ordering: ptxas turns the .release/.acquire forms into a GPU-scope MEMBAR)
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(p)), "r"(rank));
  return out;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release.cta): a .release.cluster here compiles to a GPU-scope MEMBAR (~1000 cycles)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D tiled load issued by either CTA of a pair: data lands in THIS CTA's shared memory, the bytes are counted on the
// barrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// L2 eviction-priority hints for TMA (the encodings createpolicy.fractional.L2::evict_*.b64 produces for fraction 1.0).  A
// training step streams ~2 GB of single-use data (bf16 weights, fp32 weight gradients) through a 126 MB L2 that also holds
// the ~100 MB of activations every kernel re-reads: weights and weight gradients are marked evict-first.
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                      int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* m, const void* smem_src, int c0, int c1, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d_hint(const CUtensorMap* m, const void* smem_src, int c0, int c1, uint64_t policy) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair_warp(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::"r"(
          smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// as tma_load_2d_pair, delivered to the same shared-memory offset of every CTA in `cta_mask`; each copy is counted on
// the barrier (same offset as `bar_cluster_addr`) of the receiving CTA's pair leader
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                    int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// programmatic dependent launch (see launch.h)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 32 lanes x 32 columns (fp32) -> 32 registers per thread; thread i of the warp reads TMEM lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// UMMA descriptors (sm_100 shared-memory matrix descriptor + kind::f16 instruction descriptor)
//   smem descriptor bits: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 |
//                         [61,64) layout (2 = SWIZZLE_128B)
// All operand tiles in this library use the 128-byte swizzle written by TMA (or by hand with
// swz128()), 64 bf16 along the contiguous dimension:
//   K-major  tile [rows][64 k]  : SBO = 1024 (8 rows x 128 B), LBO unused
//   MN-major tile [k rows][64 mn]: SBO = 1024 (8 k-rows x 128 B), LBO = byte distance to the next
//                                  64-wide MN chunk
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M x N tile, per-operand major-ness
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                              // D format  = F32
         | (1u << 7)                            // A format  = BF16
         | (1u << 10)                           // B format  = BF16
         | ((a_mn_major ? 1u : 0u) << 15)       // A major
         | ((b_mn_major ? 1u : 0u) << 16)       // B major
         | ((N >> 3) << 17)                     // N / 8
         | ((M >> 4) << 24);                    // M / 16
}

// byte offset of (row, 16-byte chunk) inside a 128B-swizzled tile whose rows are 128 bytes
__device__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t chunk16) {
  return row * 128u + ((chunk16 ^ (row & 7u)) << 4);
}

// ----------------------------------------------------------------------------------------------
// numerics helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Cheap erf-GELU for the GEMM epilogues: Phi(x) = 0.5 (1 + erf(x / sqrt 2)) ~= 0.5 + 0.5 tanh(x (a + b x^2 + c x^4)), a minimax fit
// (max |x Phi - x Phi_fit| = 2.5e-5 over the reals, x^2 clamped at 36 where tanh has saturated) evaluated with the hardware
// tanh (MUFU, rel. error 2^-11).  Both errors are an order of magnitude below the bf16 rounding of the stored result.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float phi_tanh_arg(float x) {
  const float u = fminf(x * x, 36.0f);
  return x * fmaf(u, fmaf(u, -3.51902393e-4f, 3.70080200e-2f), 7.97505275e-1f);
}
// two tanh per MUFU operation (tanh.approx.f16x2): the epilogues of the GELU GEMMs are bound by the MUFU pipe (16 results per
// clock per SM) once everything else is out of the way.  The f16 rounding of argument and result moves x Phi(x) by at most
// 4.6e-4 absolute (2.4e-4 |x|), an order of magnitude below the bf16 rounding of the stored value.
__device__ __forceinline__ float2 tanh2_fast(float a, float b) {
  uint32_t h, t;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));      // (hi, lo) = (b, a)
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h));
  float2 r;
  asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(r.x), "=f"(r.y) : "r"(t));
  return r;
}
// x Phi(x) for a pair
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  const float2 t = tanh2_fast(phi_tanh_arg(x0), phi_tanh_arg(x1));
  const float h0 = 0.5f * x0, h1 = 0.5f * x1;
  x0 = fmaf(h0, t.x, h0);
  x1 = fmaf(h1, t.y, h1);
}
// d/dx of the SAME fitted function 0.5 x (1 + tanh(u(x))), u = x (a + b x^2 + c x^4): 0.5 (1 + t) + 0.5 x (1 - t^2) u'(x), with
// u' = a + 3 b x^2 + 5 c x^4 (zero where x^2 is clamped: t has saturated there).  One tanh instead of tanh + exp2; max error
// against the exact erf-GELU derivative 1.1e-4 (1.7e-3 with the f16x2 tanh).
__device__ __forceinline__ void gelu_fast_grad2(float x0, float x1, float& g0, float& g1) {
  const float2 t = tanh2_fast(phi_tanh_arg(x0), phi_tanh_arg(x1));
  const float u0 = x0 * x0, u1 = x1 * x1;
  const float d0 = u0 < 36.0f ? fmaf(u0, fmaf(u0, 5.0f * -3.51902393e-4f, 3.0f * 3.70080200e-2f), 7.97505275e-1f) : 0.0f;
  const float d1 = u1 < 36.0f ? fmaf(u1, fmaf(u1, 5.0f * -3.51902393e-4f, 3.0f * 3.70080200e-2f), 7.97505275e-1f) : 0.0f;
  g0 = fmaf(0.5f * x0 * fmaf(-t.x, t.x, 1.0f), d0, fmaf(0.5f, t.x, 0.5f));
  g1 = fmaf(0.5f * x1 * fmaf(-t.y, t.y, 1.0f), d1, fmaf(0.5f, t.y, 0.5f));
}
// x Phi(x) AND its derivative for a pair, from one tanh per pair: x0 / x1 become the activation, g0 / g1 the derivative
__device__ __forceinline__ void gelu_fast2_with_grad(float& x0, float& x1, float& g0, float& g1) {
  const float2 t = tanh2_fast(phi_tanh_arg(x0), phi_tanh_arg(x1));
  const float u0 = x0 * x0, u1 = x1 * x1;
  const float d0 = u0 < 36.0f ? fmaf(u0, fmaf(u0, 5.0f * -3.51902393e-4f, 3.0f * 3.70080200e-2f), 7.97505275e-1f) : 0.0f;
  const float d1 = u1 < 36.0f ? fmaf(u1, fmaf(u1, 5.0f * -3.51902393e-4f, 3.0f * 3.70080200e-2f), 7.97505275e-1f) : 0.0f;
  const float h0 = 0.5f * x0, h1 = 0.5f * x1;
  g0 = fmaf(h0 * fmaf(-t.x, t.x, 1.0f), d0, fmaf(0.5f, t.x, 0.5f));
  g1 = fmaf(h1 * fmaf(-t.y, t.y, 1.0f), d1, fmaf(0.5f, t.y, 0.5f));
  x0 = fmaf(h0, t.x, h0);
  x1 = fmaf(h1, t.y, h1);
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_fast(phi_tanh_arg(x)), hx);
}
// d/dx [x Phi(x)] = Phi(x) + x phi(x), phi(x) = exp(-x^2/2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_fast_grad(float x) {
  const float cdf = fmaf(0.5f, tanh_fast(phi_tanh_arg(x)), 0.5f);
  const float pdf = 0.39894228040143267794f * exp2f(-0.72134752044448170368f * x * x);
  return fmaf(x, pdf, cdf);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------------------------
// Philox4x32-7 counter RNG for dropout (seven rounds: the smallest round count of the Philox family that passes BigCrush,
// Salmon et al. SC'11 table 2; the usual ten add a safety margin that a dropout mask does not need, and the rounds are a
// measurable share of the LayerNorm / attention kernels).  The mask of element `idx` of dropout site `site` is a pure
// function of (seed, site, idx), so forward and backward regenerate it instead of storing it.
// One Philox call yields 4 x 32 random bits = the keep decisions of 4 consecutive elements.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// 128 random bits = eight 16-bit keep decisions: elements [8*oct, 8*oct+8) of dropout site `site`; element j uses the
// low (j even) / high (j odd) half of word j/2.  16 bits per decision halve the Philox work; p is realised as
// round(p * 65536) / 65536 (0.1 -> 0.100006).
__device__ __forceinline__ uint4 dropout_bits(uint64_t seed, uint32_t site, uint32_t oct) {
  return philox4x32(oct, site, 0x5EEDu, 0u, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}
// 16-bit threshold such that P(bits16 >= thr) = 1 - p  (keep)
__host__ __device__ inline uint32_t dropout_threshold(float p) {
  double t = static_cast<double>(p) * 65536.0 + 0.5;
  if (t < 0.0) t = 0.0;
  if (t > 65535.0) t = 65535.0;
  return static_cast<uint32_t>(t);
}
// bit j of the result = keep decision of element 8*oct + j
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint32_t site, uint32_t oct, uint32_t thr) {
  const uint4 r = dropout_bits(seed, site, oct);
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) >= thr ? 1u : 0u) << 0; m |= ((r.x >> 16) >= thr ? 1u : 0u) << 1;
  m |= ((r.y & 0xFFFFu) >= thr ? 1u : 0u) << 2; m |= ((r.y >> 16) >= thr ? 1u : 0u) << 3;
  m |= ((r.z & 0xFFFFu) >= thr ? 1u : 0u) << 4; m |= ((r.z >> 16) >= thr ? 1u : 0u) << 5;
  m |= ((r.w & 0xFFFFu) >= thr ? 1u : 0u) << 6; m |= ((r.w >> 16) >= thr ? 1u : 0u) << 7;
  return m;
}

}  // namespace vb

// host-side helpers -------------------------------------------------------------------------------
#define VB_CUDA_CHECK(expr)                                  \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) {                                 \
      vb_set_last_error(#expr, cudaGetErrorString(_e));      \
      return VB_ERR_CUDA;                                \
    }                                                        \
  } while (0)

#define VB_REQUIRE(cond, msg)                 \
  do {                                        \
    if (!(cond)) {                            \
      vb_set_last_error(#cond, msg);          \
      return VB_ERR_BAD_ARG;              \
    }                                         \
  } while (0)

extern "C" void vb_set_last_error(const char* what, const char* detail);

#define VB_STR(x) #x
#define VB_STR2(x) VB_STR(x)
