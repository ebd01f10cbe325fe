// tcgen05 / TMEM GEMM for sm_100a, fed by TMA.  One persistent, warp-specialised kernel:
//
//   warp 0      TMA producer      global -> 128B-swizzled smem ring (mbarrier full/empty)
//   warp 1      MMA issuer        one thread issues tcgen05.mma (bf16 -> fp32 in TMEM)
//   warps 2..9  epilogue          tcgen05.ld accumulator -> scale/bias/aux/activation -> swizzled smem -> TMA store
//
// CG = 2 (the normal case): the kernel runs as CTA PAIRS (2-wide clusters = the two SMs of a TPC) and issues
// tcgen05.mma.cta_group::2 with M = 256: each CTA holds its own 128 rows of A, HALF of the B tile and its own 128
// accumulator rows.  At the ViLBERT shapes (M = 1600 / 2048) the main loop is bound by the bytes an SM can pull from L2
// (~50-58 B/clk/SM measured, 64 nominal), not by the tensor pipe, so halving the B bytes per SM is what moves the needle:
// a 256 x BN pair tile ingests (16 KB + BN*64 B) per SM per 64-deep k-block.  The tile width BN is a RUN-TIME parameter
// (any multiple of 32 up to 256: UMMA N is free in steps of 16) picked per problem so that the pair tiles fill the 74 TPCs
// in as few waves as possible.  CG = 1 (single CTA, M = 128) remains for one-row-block problems (poolers, classifier).
//
// Round-2 structure (what the in-kernel clock64 timeline of round 1 asked for; the round-2 timeline is profiles/r02_gemm_timeline.txt):
//   * prologue: every thread arrives on the cluster barrier at entry; barrier init, TMEM allocation and descriptor prefetch
//     run in parallel behind it, and the producer issues its first loads as soon as the barrier and griddepcontrol.wait
//     return (first load ~1800 -> ~700 cycles after entry).
//   * epilogue: WARP-AUTONOMOUS.  Each epilogue warp owns 32 accumulator rows x the even or odd 32-column chunks of the
//     tile, a private staging box and its own TMA stores: no CTA-wide barrier, no shared staging panel that has to drain
//     before the next one is filled.  Bias / scale go to a warp-private smem strip while the main loop runs; the aux operand
//     (residual gradient or GELU pre-activation) is read straight from global memory into registers one chunk ahead, so it
//     costs no shared memory (dgrad GEMMs get the same ring depth as forward ones) and its latency hides behind the main loop.
//
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.  Operands may be K-major or
// MN-major (descriptor + TMA box change only), so forward, dgrad and wgrad of nn.Linear (reference
// models/vilbert_facebook_arch.py:127-129 etc.) all run here without a transposed copy of anything.  The kernel is
// launched with programmatic dependent launch.  See include/vilbert_b200.h for the ABI.
//
// CONV kernels (implicit-GEMM convolution, the 3x3 and strided 1x1 layers of the ResNet trunks): the A operand is the NHWC
// activation itself, described by an IM2COL tensor map (cuTensorMapEncodeIm2col).  A k-block is one filter tap x 64 input
// channels; the producer decodes the tile's first output pixel once per tile, steps (channel block, kx, ky) counters per
// k-block and issues cp.async.bulk.tensor.4d...im2col with the tap as the instruction's offset operand.  The bytes land in
// shared memory exactly as a K-major [128 pixels, 64] box would (zero padding = out-of-bounds fill), so the MMA issuer, the
// multicast halves and the epilogue are untouched and the result is bit-identical to the GEMM over a materialised
// [pixels, kh*kw*Cin] matrix -- which no longer exists (it was 38 % of the RoI stage's kernel time, profiles/r02_roi_stage.md).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "launch.h"
#include "../../include/vilbert_b200.h"
#include "tensormap.h"

namespace vb {

constexpr int GEMM_BM = 128;        // output rows owned by one CTA (a pair covers 256)
constexpr int GEMM_BK = 64;         // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_EPI_WARPS = 8;   // four TMEM sub-partitions x (even | odd 32-column chunks)
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_ACC_STAGES = 2;
constexpr int GEMM_MAX_STAGES = 10;
constexpr int GEMM_CHUNK = 32;            // epilogue granule: 32 accumulator columns = one staging box per warp
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;
constexpr int GEMM_BOX_BF16 = 32 * GEMM_CHUNK * 2;   // 32 rows x 64 B, 64B-swizzled (one per epilogue warp)
constexpr int GEMM_BOX_F32 = 32 * GEMM_CHUNK * 4;    // 32 rows x 128 B, 128B-swizzled
constexpr int GEMM_NBARS = 2 * GEMM_MAX_STAGES + 2 * GEMM_ACC_STAGES;
#ifdef VB_GEMM_TRACE
constexpr int GEMM_SMEM_SLACK = 768;   // static shared memory: barriers + the trace buffer
#else
constexpr int GEMM_SMEM_SLACK = 512;   // static shared memory: barriers + TMEM base slot
#endif
// Tiles up to 128 columns wide use a compact configuration (<= 112 KB smem, <= 102 registers, <= 256 TMEM columns) so that
// TWO CTAs (of different kernels: the text / visual / weight-gradient streams) can share an SM and overlap each other's
// prologue and epilogue.  VB_GEMM_OCC1=1 switches it off (experiments).
__host__ __device__ constexpr int gemm_smem_limit(int occ) { return (occ == 2 ? 112 * 1024 : 227 * 1024) - GEMM_SMEM_SLACK; }

// shared-memory carve-up (offsets from the 1024-byte aligned base), the same arithmetic on host and device
struct GemmSmem { uint32_t b, out, x, aux, bias, scale, total; };
// nbuf = staging boxes per epilogue warp (2: a chunk is staged while the previous one is still being read by its TMA store)
__host__ __device__ inline GemmSmem gemm_smem(int stages, int bnl, int bn, bool f32, bool preact, bool has_scale, bool has_aux, int nbuf) {
  GemmSmem s;
  s.b = static_cast<uint32_t>(stages) * GEMM_A_BYTES;
  s.out = s.b + static_cast<uint32_t>(stages * bnl) * 128u;
  s.x = s.out + GEMM_EPI_WARPS * nbuf * (f32 ? GEMM_BOX_F32 : GEMM_BOX_BF16);
  s.aux = s.x + (preact ? GEMM_EPI_WARPS * nbuf * GEMM_BOX_BF16 : 0);
  s.bias = s.aux + (has_aux ? GEMM_EPI_WARPS * 2 * GEMM_BOX_BF16 : 0);      // aux boxes are always double-buffered
  const uint32_t strip = static_cast<uint32_t>((bn / GEMM_CHUNK + 1) / 2) * GEMM_CHUNK * 4u;   // floats of one warp's chunks
  s.scale = s.bias + GEMM_EPI_WARPS * strip;
  s.total = s.scale + (has_scale ? GEMM_EPI_WARPS * strip : 0u);
  return s;
}

struct GemmKernelParams {
  const float* scale;
  const float* bias;
  const __nv_bfloat16* aux;   // read straight from global memory into registers by the epilogue warps
  long long ld_aux;
  int m, n, k;
  int bn;           // tile width (columns of the output per CTA pair / per lone CTA)
  int d_is_f32, reduce_add, act, aux_mode, has_preact, preact_grad;
  int splits, kb_per_split;
  int m_tiles, n_tiles;
  int stages;       // depth of the operand ring
  int nbuf;         // staging boxes per epilogue warp (1 | 2)
  int pdl_late;     // trigger the dependent launch when the producer is done instead of at kernel entry
  int tmem_cols;
  int debug_mode;   // profiling only (VB_GEMM_DEBUG): 1 = no MMA issue, 2 = no TMA loads; results are garbage
  uint32_t magic_m, magic_mn;   // fast_div multipliers for m_tiles and m_tiles * n_tiles
  unsigned long long b_policy, d_policy;   // L2 eviction hints for the B loads / D stores
  long long* trace; // profiling only (vb_gemm_set_trace): clock64 stamps per CTA, NULL in production
  // implicit-GEMM convolution (CONV kernels): A is an NHWC activation read through an im2col tensor map
  int conv_hw_out, conv_w_out;     // output pixels per image / per row
  int conv_stride, conv_pad, conv_kw;
  int conv_cblocks;                // 64-channel blocks per filter tap (c / 64)
};

// compiled in only with -DVB_GEMM_TRACE (__graft_entry__.build_variant builds that copy of the library): even a
// predicated-off stamp is instructions on the cold, instruction-fetch-bound path of a 5 us kernel
#ifdef VB_GEMM_TRACE
__shared__ long long s_trace[24];   // stamps go to shared memory and are dumped at exit: global stores would perturb
#endif
__device__ __forceinline__ void trace_stamp(int slot) {
#ifdef VB_GEMM_TRACE
  s_trace[slot] = clock64();
#endif
}

// Kernel parameters live in the constant bank, and ptxas re-loads them at every use because such loads are "free" -- but a
// cold LDC / LDCU costs ~100 cycles, and a 5 us kernel whose epilogue tests four flags per chunk pays that latency in
// a dependent chain, dozens of times.  pin() forces a value into a register once, at kernel entry, where all the loads
// overlap each other and the TMEM allocation.
__device__ __forceinline__ void pin(int& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void pin(uint32_t& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void pin(long long& x) { asm volatile("" : "+l"(x)); }
template <typename T>
__device__ __forceinline__ void pin(T*& x) { asm volatile("" : "+l"(x)); }

// shared-space accessors on 32-bit shared addresses (the staging pointers are carved from dynamic smem at run time, so
// plain C++ dereferences would compile to generic LD.E / ST.E)
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
// CTA barrier 0, callable from the (warp-uniform) role branches: every thread of the CTA executes exactly one of them
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }
// named barrier 1: the MMA warp (publishes the TMEM base address) + the epilogue warps (consume it)
__device__ __forceinline__ void tmem_slot_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(32 + GEMM_EPI_WARPS * 32) : "memory"); }

// Warp-collective forms for the producer: the WHOLE (converged) warp executes the call with warp-uniform operands and one
// elected lane issues.  With `if (lane == 0) tma_load(...)` the operands live in per-lane registers and ptxas wraps every
// UTMALDG in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop; the producer thread then spends ~500 cycles of dependent scalar
// instructions per k-block (measured with tools/gemm_mainloop.py: the k-block time was ~520 cycles whatever the tile width,
// the ring depth or whether any MMA was issued) -- more than the 256-cycle MMA time of a 128-wide tile.
__device__ __forceinline__ void mbar_arrive_expect_tx_warp(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_warp(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint64_t policy) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
               " [%0], [%1, {%3, %4}], [%2], %5;\n\t}" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_mc_warp(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
               " [%0], [%1, {%3, %4}], [%2], %5;\n\t}" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_warp(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint64_t policy) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
               " [%0], [%1, {%3, %4}], [%2], %5;\n\t}" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
}
// im2col-mode loads (implicit-GEMM convolution): `pixels` consecutive output pixels x 64 channels of filter tap (ox, oy),
// starting at the output pixel whose base input position is (w, h) of image n; same shared-memory layout as a K-major box
__device__ __forceinline__ void tma_load_im2col_warp(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n,
                                                     uint16_t ox, uint16_t oy) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};\n\t}" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c),
               "r"(w), "r"(h), "r"(n), "h"(ox), "h"(oy)
               : "memory");
}
__device__ __forceinline__ void tma_load_im2col_pair_warp(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n,
                                                          uint16_t ox, uint16_t oy) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};\n\t}" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c),
               "r"(w), "r"(h), "r"(n), "h"(ox), "h"(oy)
               : "memory");
}
__device__ __forceinline__ void tma_load_im2col_pair_mc_warp(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h,
                                                             int n, uint16_t ox, uint16_t oy, uint16_t mask) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes.multicast::cluster"
               " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;\n\t}" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar),
               "r"(c), "r"(w), "r"(h), "r"(n), "h"(ox), "h"(oy), "h"(mask)
               : "memory");
}
// every lane polls (same barrier, same answer: no divergence), so the code after the wait is still warp-uniform
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    if ((++spins & 1023u) == 0) {      // bounded by WALL CLOCK (2 s): see mbar_wait
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) __trap();
    }
  }
}

// Epilogue kinds: the common ones are compiled as separate kernels so that a chunk's epilogue is ~100 straight-line
// instructions with 32-wide instruction-level parallelism and no flag tests (two epilogue warps per scheduler cannot hide
// branch / LDS latencies, and the code runs once per launch from a cold instruction cache); everything else takes the generic
// kernel with run-time flags.
// EPI_BN_RELU / EPI_BN_ADD_RELU: the ResNet convolutions (folded-BatchNorm scale + bias, [+ bottleneck identity,] ReLU, bf16 out)
enum { EPI_GENERIC = 0, EPI_BIAS = 1, EPI_GELU_PRE = 2, EPI_AUX_ADD = 3, EPI_AUX_GELUGRAD = 4, EPI_F32 = 5, EPI_BN_RELU = 6, EPI_BN_ADD_RELU = 7,
       EPI_GELU_GRADPRE = 8,   // bias + GELU, second output = GELU'(pre-activation) (what the backward multiplies by)
       EPI_AUX_MUL = 9 };      // acc * aux (the dgrad of a GELU layer whose forward stored GELU')

// x / d for x * d < 2^32, d >= 1, with magic = ceil(2^32 / d) computed on the host (d = 1 -> magic 0 = "identity")
__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t magic) { return magic == 0u ? x : __umulhi(x, magic); }

struct TileCoord { int m_idx, n_idx, split; };
__device__ __forceinline__ TileCoord decode_tile(const GemmKernelParams& p, int tile) {
  const uint32_t split = fast_div(static_cast<uint32_t>(tile), p.magic_mn);
  const uint32_t mn = static_cast<uint32_t>(tile) - split * static_cast<uint32_t>(p.m_tiles * p.n_tiles);
  const uint32_t n_idx = fast_div(mn, p.magic_m);
  return {static_cast<int>(mn - n_idx * static_cast<uint32_t>(p.m_tiles)), static_cast<int>(n_idx), static_cast<int>(split)};
}

// Eight columns of the epilogue: v = acc * scale + bias (+ aux | * gelu'(aux)) -> activation -> one 16-byte (bf16) or two
// 16-byte (fp32) pieces of the warp's staging box.  `g` = which eighth-of-a-chunk (0..3), a compile-time constant at every
// call site (the chunk is unrolled).
template <int EPI>
__device__ __forceinline__ void epilogue_octet(const GemmKernelParams& p, const uint32_t* r /* 8 accumulator words */, int g,
                                               uint32_t s_bias, uint32_t s_scale, uint32_t arow, uint32_t xrow, uint32_t orow,
                                               uint32_t swz64, uint32_t swz128) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
  if constexpr (EPI == EPI_BN_RELU || EPI == EPI_BN_ADD_RELU) {
    const float4 s0 = lds_f4(s_scale + static_cast<uint32_t>(g) * 32u), s1 = lds_f4(s_scale + static_cast<uint32_t>(g) * 32u + 16u);
    v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
  }
  if constexpr (EPI == EPI_GENERIC) {
    if (s_scale != 0u) {
      const float4 s0 = lds_f4(s_scale + static_cast<uint32_t>(g) * 32u), s1 = lds_f4(s_scale + static_cast<uint32_t>(g) * 32u + 16u);
      v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
    }
  }
  if constexpr (EPI == EPI_GENERIC || EPI == EPI_BIAS || EPI == EPI_GELU_PRE || EPI == EPI_GELU_GRADPRE || EPI == EPI_BN_RELU ||
                EPI == EPI_BN_ADD_RELU) {
    const float4 b0 = lds_f4(s_bias + static_cast<uint32_t>(g) * 32u), b1 = lds_f4(s_bias + static_cast<uint32_t>(g) * 32u + 16u);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (EPI == EPI_GELU_PRE || (EPI == EPI_GENERIC && p.has_preact && !p.preact_grad))
    sts_u4(xrow + ((static_cast<uint32_t>(g) ^ swz64) << 4), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
           pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  if (EPI == EPI_GELU_GRADPRE || (EPI == EPI_GENERIC && p.has_preact && p.preact_grad)) {
    // GELU and its derivative from ONE tanh per pair: the forward has the tanh anyway, the backward then only multiplies
    float gr[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) gelu_fast2_with_grad(v[i], v[i + 1], gr[i], gr[i + 1]);
    sts_u4(xrow + ((static_cast<uint32_t>(g) ^ swz64) << 4), pack_bf16x2(gr[0], gr[1]), pack_bf16x2(gr[2], gr[3]),
           pack_bf16x2(gr[4], gr[5]), pack_bf16x2(gr[6], gr[7]));
  }
  if (EPI == EPI_AUX_ADD || EPI == EPI_AUX_GELUGRAD || EPI == EPI_AUX_MUL || EPI == EPI_BN_ADD_RELU ||
      (EPI == EPI_GENERIC && p.aux_mode != VB_AUX_NONE)) {
    // my row of the warp's aux box (TMA-loaded, 64-byte rows, 64B swizzle): piece g
    const uint4 aux = lds_u4(arow + ((static_cast<uint32_t>(g) ^ swz64) << 4));
    const float2 a0 = unpack_bf16x2(aux.x), a1 = unpack_bf16x2(aux.y), a2 = unpack_bf16x2(aux.z), a3 = unpack_bf16x2(aux.w);
    const float av[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
    if (EPI == EPI_AUX_ADD || EPI == EPI_BN_ADD_RELU || (EPI == EPI_GENERIC && p.aux_mode == VB_AUX_ADD)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += av[i];
    } else if (EPI == EPI_AUX_MUL || (EPI == EPI_GENERIC && p.aux_mode == VB_AUX_MUL)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= av[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        float g0, g1;
        gelu_fast_grad2(av[i], av[i + 1], g0, g1);
        v[i] *= g0;
        v[i + 1] *= g1;
      }
    }
  }
  if (EPI == EPI_GELU_GRADPRE || (EPI == EPI_GENERIC && p.has_preact && p.preact_grad && p.act == VB_ACT_GELU)) {
    // (the activation itself was produced together with its derivative above)
  } else if (EPI == EPI_GELU_PRE || (EPI == EPI_GENERIC && p.act == VB_ACT_GELU)) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) gelu_fast2(v[i], v[i + 1]);
  } else if (EPI == EPI_BN_RELU || EPI == EPI_BN_ADD_RELU || (EPI == EPI_GENERIC && p.act == VB_ACT_RELU)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (EPI == EPI_GENERIC && p.act == VB_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = tanh_fast(v[i]);
  }
  if (EPI == EPI_F32 || (EPI == EPI_GENERIC && p.d_is_f32)) {
    // fp32 box: 32 columns = eight 16-byte pieces per 128-byte row (128B-swizzled); this octet is pieces 2g, 2g + 1
    sts_u4(orow + ((static_cast<uint32_t>(2 * g) ^ swz128) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
           __float_as_uint(v[3]));
    sts_u4(orow + ((static_cast<uint32_t>(2 * g + 1) ^ swz128) << 4), __float_as_uint(v[4]), __float_as_uint(v[5]),
           __float_as_uint(v[6]), __float_as_uint(v[7]));
  } else {
    // bf16 box: 32 columns = four 16-byte pieces per 64-byte row (64B-swizzled); this octet is piece g
    sts_u4(orow + ((static_cast<uint32_t>(g) ^ swz64) << 4), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
           pack_bf16x2(v[6], v[7]));
  }
}

// One k-block of operand loads (the producer's loop body), executed by the whole producer warp with uniform operands.
template <bool A_MN, bool B_MN, int CG, int NP>
__device__ __forceinline__ void issue_kblock(const CUtensorMap* tma_a, const CUtensorMap* tma_b, uint32_t sa, uint32_t sb,
                                             uint32_t full_bar, uint32_t bar_leader, int bnl, int k0, int m0, int n0,
                                             uint32_t prank, uint32_t pair, uint16_t a_mask, unsigned long long b_policy) {
  const int b_bytes = bnl * GEMM_BK * 2;
  if constexpr (CG == 2) {
    // both CTAs' bytes are counted on the leader's barrier; a peer load that lands before the leader's expect_tx only
    // drives the transaction count negative for a moment (same phase: the peer cannot run ahead of the leader's MMA,
    // which frees the stage for both)
    if (prank == 0) mbar_arrive_expect_tx_warp(full_bar, static_cast<uint32_t>(2 * (GEMM_A_BYTES + b_bytes)));
    if constexpr (NP == 2) {
      // the two pairs of the cluster work on the same 256 rows: each CTA fetches HALF of its A tile (64 rows) and
      // multicasts it to itself and to its twin in the other pair, which halves the A bytes read from L2.  Safe:
      // empty_bar counts the commits of BOTH pairs, so the twin's slot is free too.
      if constexpr (A_MN) tma_load_2d_pair_mc_warp(sa + pair * (GEMM_BK * 128), tma_a, bar_leader, m0 + static_cast<int>(pair) * 64, k0, a_mask);
      else                tma_load_2d_pair_mc_warp(sa + pair * (64 * 128), tma_a, bar_leader, k0, m0 + static_cast<int>(pair) * 64, a_mask);
    } else if constexpr (A_MN) {
#pragma unroll
      for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d_pair_warp(sa + j * (GEMM_BK * 128), tma_a, bar_leader, m0 + j * 64, k0, L2_EVICT_NORMAL);
    } else {
      tma_load_2d_pair_warp(sa, tma_a, bar_leader, k0, m0, L2_EVICT_NORMAL);
    }
    if constexpr (B_MN) {
      for (int j = 0; j < bnl / 64; ++j) tma_load_2d_pair_warp(sb + j * (GEMM_BK * 128), tma_b, bar_leader, n0 + j * 64, k0, b_policy);
    } else {
      tma_load_2d_pair_warp(sb, tma_b, bar_leader, k0, n0, b_policy);
    }
  } else {
    mbar_arrive_expect_tx_warp(full_bar, static_cast<uint32_t>(GEMM_A_BYTES + b_bytes));
    if constexpr (A_MN) {
#pragma unroll
      for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d_warp(sa + j * (GEMM_BK * 128), tma_a, full_bar, m0 + j * 64, k0, L2_EVICT_NORMAL);
    } else {
      tma_load_2d_warp(sa, tma_a, full_bar, k0, m0, L2_EVICT_NORMAL);
    }
    if constexpr (B_MN) {
      for (int j = 0; j < bnl / 64; ++j) tma_load_2d_warp(sb + j * (GEMM_BK * 128), tma_b, full_bar, n0 + j * 64, k0, b_policy);
    } else {
      tma_load_2d_warp(sb, tma_b, full_bar, k0, n0, b_policy);
    }
  }
}

// The same for an implicit-GEMM convolution: A = one filter tap (kx, ky) x 64 channels [c0, c0 + 64) of this CTA's output
// pixels, fetched in im2col mode from the NHWC activation (base input position (bw, bh) of image bn_img); B K-major as above.
template <int CG, int NP>
__device__ __forceinline__ void issue_kblock_conv(const CUtensorMap* tma_a, const CUtensorMap* tma_b, uint32_t sa, uint32_t sb,
                                                  uint32_t full_bar, uint32_t bar_leader, int bnl, int k0, int n0, uint32_t prank,
                                                  uint32_t pair, uint16_t a_mask, unsigned long long b_policy, int c0, int bw, int bh,
                                                  int img, int kx, int ky) {
  const int b_bytes = bnl * GEMM_BK * 2;
  const uint16_t ox = static_cast<uint16_t>(kx), oy = static_cast<uint16_t>(ky);
  if constexpr (CG == 2) {
    if (prank == 0) mbar_arrive_expect_tx_warp(full_bar, static_cast<uint32_t>(2 * (GEMM_A_BYTES + b_bytes)));
    if constexpr (NP == 2) tma_load_im2col_pair_mc_warp(sa + pair * (64 * 128), tma_a, bar_leader, c0, bw, bh, img, ox, oy, a_mask);
    else                   tma_load_im2col_pair_warp(sa, tma_a, bar_leader, c0, bw, bh, img, ox, oy);
    tma_load_2d_pair_warp(sb, tma_b, bar_leader, k0, n0, b_policy);
  } else {
    mbar_arrive_expect_tx_warp(full_bar, static_cast<uint32_t>(GEMM_A_BYTES + b_bytes));
    tma_load_im2col_warp(sa, tma_a, full_bar, c0, bw, bh, img, ox, oy);
    tma_load_2d_warp(sb, tma_b, full_bar, k0, n0, b_policy);
  }
}

template <bool A_MN, bool B_MN, int CG, int NP, int OCC, int EPI, bool CONV = false>
__global__ void __launch_bounds__(GEMM_THREADS, OCC)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_x,
                 const __grid_constant__ CUtensorMap tma_aux, const GemmKernelParams p_const) {
  constexpr int CS = CG * NP;                            // CTAs per cluster: NP pairs, side by side along N, sharing A
  static_assert(CG == 2 || NP == 1, "multicast clusters are built from CTA pairs");
  // barriers live in STATIC shared memory: their addresses are link-time constants, so warp 0 can initialise them with its very
  // first instructions -- before a single kernel parameter has been read -- and the cluster barrier that publishes them
  // completes that much earlier
  __shared__ __align__(8) uint64_t bars[GEMM_NBARS];
  __shared__ __align__(8) uint64_t aux_bars[GEMM_EPI_WARPS * 2];     // one per (epilogue warp, aux box)
  __shared__ uint32_t tmem_base_slot[2];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { trace_stamp(0); tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b); }
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + GEMM_MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * GEMM_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + GEMM_ACC_STAGES;

  // ------------------------------------------------------------------ prologue
  // warp 0 initialises the barriers and publishes them to the cluster (fence.mbarrier_init.release + relaxed arrive, the
  // peers' barrier.cluster.wait is the acquire); every other warp arrives at once, so the cluster barrier completes as soon
  // as the slowest warp 0 of the cluster is done.  TMEM allocation and descriptor prefetch run behind the arrive.
  if (warp == 0) {
    // one barrier per lane, branch-free: [0,10) full (1: the leader's producer arrives, expecting the bytes of BOTH CTAs),
    // [10,20) empty (one multicast commit per pair of the cluster), 20/21 tmem_full (1), 22/23 tmem_empty (one arrive per
    // epilogue warp of every CTA of the pair)
    const bool is_tmem_empty = lane >= 2 * GEMM_MAX_STAGES + GEMM_ACC_STAGES;
    const bool is_empty = lane >= GEMM_MAX_STAGES && lane < 2 * GEMM_MAX_STAGES;
    if (lane < GEMM_NBARS) mbar_init(&bars[lane], is_tmem_empty ? GEMM_EPI_WARPS * CG : (is_empty ? NP : 1));
    if (lane < GEMM_EPI_WARPS * 2) mbar_init(&aux_bars[lane], 1);
    __syncwarp();
    if (lane == 0) {
      fence_mbar_init();
      trace_stamp(19);
    }
    __syncwarp();
  }
  if constexpr (CS > 1) cluster_arrive_relaxed();
  if (warp == 2 && lane == 0) tma_prefetch_desc(&tma_d);
#ifdef VB_GEMM_TRACE
  if (threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); s_trace[22] = static_cast<long long>(g); }
#endif

  GemmKernelParams p = p_const;
  pin(p.bn); pin(p.stages); pin(p.d_is_f32); pin(p.has_preact); pin(p.nbuf); pin(p.pdl_late);
  // PDL: the next kernel of the stream may begin its own prologue; it blocks in griddepcontrol.wait until this grid is done.
  // pdl_late (default; VB_GEMM_PDL_LATE=0 for the early form): the producer triggers the dependent launch after its LAST load
  // instead of at kernel entry -- a dependent that has sat long in griddepcontrol.wait wakes up late (4.73 -> 4.68 ms/step).
  if (!p.pdl_late) griddep_launch();
  pin(p.scale); pin(p.bias); pin(p.aux); pin(p.ld_aux); pin(p.m); pin(p.n); pin(p.k);
  pin(p.reduce_add); pin(p.act); pin(p.aux_mode);
  pin(p.splits); pin(p.kb_per_split); pin(p.m_tiles); pin(p.n_tiles); pin(p.tmem_cols);
  pin(p.magic_m); pin(p.magic_mn);
  asm volatile("" : "+l"(p.b_policy)); asm volatile("" : "+l"(p.d_policy));
  const int BN = p.bn;
  const int BNL = BN / CG;                               // B rows (columns of the output) this CTA loads
  const int B_BYTES = BNL * GEMM_BK * 2;
  const int STAGES = p.stages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const GemmSmem L = gemm_smem(STAGES, BNL, BN, p.d_is_f32 != 0, p.has_preact != 0, p.scale != nullptr, p.aux_mode != VB_AUX_NONE, p.nbuf);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + L.b;

  const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
  const uint32_t prank = rank & static_cast<uint32_t>(CG - 1);   // rank inside the pair: 0 = leader (issues the MMAs)
  const uint32_t pair = CG == 2 ? rank >> 1 : 0u;                // which pair of the cluster
  const uint32_t leader_rank = rank - prank;

  const int total_kb = (p.k + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = p.m_tiles * p.n_tiles * p.splits;
  const int first_tile = blockIdx.x / CS;
  const int tile_step = gridDim.x / CS;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one per CTA; warp-uniform code)
    const uint32_t full_leader = CG == 2 ? mapa_u32(&full_bar[0], leader_rank) : 0u;
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
    const uint16_t a_mask = static_cast<uint16_t>((1u << prank) | (1u << (prank + 2)));   // me and my twin in the other pair
    const unsigned long long b_policy = p.b_policy ? p.b_policy : L2_EVICT_NORMAL;
    if constexpr (CS > 1) cluster_wait(); else cta_sync();
    if (lane == 0) trace_stamp(1);
    griddep_wait();     // operands written by the previous kernel of the stream are complete and visible from here on
    if (lane == 0) trace_stamp(2);
    int stage = 0, fills = 0;
    uint32_t phase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile(p, tile);
      const int m0 = tc.m_idx * (GEMM_BM * CG) + static_cast<int>(prank) * GEMM_BM;
      const int n0 = (tc.n_idx * NP + static_cast<int>(pair)) * BN + static_cast<int>(prank) * BNL;
      const int kb0 = tc.split * p.kb_per_split;
      const int kb1 = min(total_kb, kb0 + p.kb_per_split);
      // implicit-GEMM convolution: where this CTA's first output pixel sits (once per tile), then (tap, channel block) counters
      int cv_img = 0, cv_bw = 0, cv_bh = 0, cv_c = 0, cv_kx = 0, cv_ky = 0;
      if constexpr (CONV) {
        const int pix = m0 + (NP == 2 ? static_cast<int>(pair) * 64 : 0);
        cv_img = pix / p.conv_hw_out;
        const int r = pix - cv_img * p.conv_hw_out;
        const int oy = r / p.conv_w_out;
        cv_bh = oy * p.conv_stride - p.conv_pad;
        cv_bw = (r - oy * p.conv_w_out) * p.conv_stride - p.conv_pad;
        const int tap = kb0 / p.conv_cblocks;
        cv_c = kb0 - tap * p.conv_cblocks;
        cv_ky = tap / p.conv_kw;
        cv_kx = tap - cv_ky * p.conv_kw;
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        if (fills >= STAGES) mbar_wait_addr(empty0 + static_cast<uint32_t>(stage) * 8u, phase ^ 1u);   // the first pass over the ring needs no wait
        ++fills;
#ifdef VB_GEMM_TRACE
        if (p.debug_mode == 2) {
          if (prank == 0 && lane == 0) mbar_arrive(&full_bar[stage]);
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          continue;
        }
#endif
        if constexpr (CONV) {
          issue_kblock_conv<CG, NP>(&tma_a, &tma_b, sa0 + static_cast<uint32_t>(stage * GEMM_A_BYTES),
                                    sb0 + static_cast<uint32_t>(stage * B_BYTES), full0 + static_cast<uint32_t>(stage) * 8u,
                                    full_leader + static_cast<uint32_t>(stage) * 8u, BNL, kb * GEMM_BK, n0, prank, pair, a_mask,
                                    b_policy, cv_c * GEMM_BK, cv_bw, cv_bh, cv_img, cv_kx, cv_ky);
          if (++cv_c == p.conv_cblocks) { cv_c = 0; if (++cv_kx == p.conv_kw) { cv_kx = 0; ++cv_ky; } }
        } else {
          issue_kblock<A_MN, B_MN, CG, NP>(&tma_a, &tma_b, sa0 + static_cast<uint32_t>(stage * GEMM_A_BYTES),
                                           sb0 + static_cast<uint32_t>(stage * B_BYTES), full0 + static_cast<uint32_t>(stage) * 8u,
                                           full_leader + static_cast<uint32_t>(stage) * 8u, BNL, kb * GEMM_BK, m0, n0, prank, pair,
                                           a_mask, b_policy);
        }
        if (lane == 0 && tile == first_tile && kb == kb0) trace_stamp(3);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (lane == 0) trace_stamp(4);
    if (p.pdl_late) griddep_launch();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ TMEM owner + MMA issuer (the leader CTA of every pair)
    if constexpr (CG == 2) {
      tmem_alloc_pair(tmem_base_slot, static_cast<uint32_t>(p.tmem_cols));
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_base_slot, static_cast<uint32_t>(p.tmem_cols));
      tmem_relinquish();
    }
    tc_fence_before();
    tmem_slot_barrier();        // publishes the TMEM base address to the epilogue warps
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot[0];
    if (lane == 0) trace_stamp(20);
    if constexpr (CS > 1) cluster_wait(); else cta_sync();
    if (prank == 0) {
      const uint32_t idesc = umma_idesc_bf16(GEMM_BM * CG, static_cast<uint32_t>(BN), A_MN, B_MN);
      const uint16_t all_mask = static_cast<uint16_t>((1u << CS) - 1u);            // smem slots are freed for the whole cluster
      const uint16_t pair_mask = static_cast<uint16_t>(3u << (pair * 2));          // accumulators are per pair
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0, tiles_done = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int split = static_cast<int>(fast_div(static_cast<uint32_t>(tile), p.magic_mn));
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        if (tiles_done >= GEMM_ACC_STAGES) {                      // both accumulator stages start out free
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
          tc_fence_after();
        }
        ++tiles_done;
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          // descriptors differ from the stage base only in the 14-bit start-address field: +32 B per K step inside a
          // K-major swizzle row, +2048 B (16 k-rows) per K step of an MN-major tile (built BEFORE the wait: off the critical path)
          const uint64_t da0 = A_MN ? umma_smem_desc(smem_u32(smem_a + stage * GEMM_A_BYTES), GEMM_BK * 128, 1024)
                                    : umma_smem_desc(smem_u32(smem_a + stage * GEMM_A_BYTES), 16, 1024);
          const uint64_t db0 = B_MN ? umma_smem_desc(smem_u32(smem_b + stage * B_BYTES), GEMM_BK * 128, 1024)
                                    : umma_smem_desc(smem_u32(smem_b + stage * B_BYTES), 16, 1024);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (tile == first_tile && kb == kb0 && lane == 0) trace_stamp(5);
#ifdef VB_GEMM_TRACE
          if (p.debug_mode == 1) {
            if constexpr (CG == 2) umma_commit_pair_warp(&empty_bar[stage], all_mask); else umma_commit_warp(&empty_bar[stage]);
          } else
#endif
          {
#pragma unroll
            for (int kk = 0; kk < GEMM_BK / 16; ++kk) {
              const uint64_t da = da0 + static_cast<uint64_t>(kk * (A_MN ? 128 : 2));
              const uint64_t db = db0 + static_cast<uint64_t>(kk * (B_MN ? 128 : 2));
              const uint32_t accum = (kb > kb0 || kk > 0) ? 1u : 0u;
              if constexpr (CG == 2) umma_bf16_pair_warp(tmem_d, da, db, idesc, accum);
              else                   umma_bf16_warp(tmem_d, da, db, idesc, accum);
            }
            // smem slot reusable (in both CTAs) once these MMAs retire
            if constexpr (CG == 2) umma_commit_pair_warp(&empty_bar[stage], all_mask); else umma_commit_warp(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if constexpr (CG == 2) umma_commit_pair_warp(&tmem_full_bar[acc], pair_mask); else umma_commit_warp(&tmem_full_bar[acc]);
        __syncwarp();
        if (++acc == GEMM_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
      if (lane == 0) trace_stamp(6);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9 of every CTA), warp-autonomous
    const int ewarp = warp - 2;
    const int quarter = warp & 3;           // TMEM sub-partition this warp may read: lanes [32q, 32q+32)
    const int half = ewarp >> 2;            // this warp owns chunks half, half + 2, half + 4, ... of every tile
    const bool tracer = ewarp == 0 && lane == 0;
    const int row = quarter * 32 + lane;    // row inside this CTA's 128-row block
    const uint32_t swz128 = static_cast<uint32_t>(lane & 7);          // 128-byte rows (fp32 boxes)
    const uint32_t swz64 = static_cast<uint32_t>((lane >> 1) & 3);    // 64-byte rows (bf16 boxes)
    const bool has_aux = p.aux_mode != VB_AUX_NONE;
    const uint32_t strip = static_cast<uint32_t>((BN / GEMM_CHUNK + 1) / 2) * GEMM_CHUNK * 4u;
    const uint32_t sa_bias = smem_u32(smem + L.bias) + static_cast<uint32_t>(ewarp) * strip;
    const uint32_t sa_scale = p.scale != nullptr ? smem_u32(smem + L.scale) + static_cast<uint32_t>(ewarp) * strip : 0u;
    const uint32_t out_box_bytes = p.d_is_f32 ? GEMM_BOX_F32 : GEMM_BOX_BF16;
    // my staging boxes (nbuf of each kind, used alternately) and my two aux boxes
    const uint32_t box_out0 = smem_u32(smem + L.out) + static_cast<uint32_t>(ewarp * p.nbuf) * out_box_bytes;
    const uint32_t box_x0 = smem_u32(smem + L.x) + static_cast<uint32_t>(ewarp * p.nbuf) * GEMM_BOX_BF16;
    const uint32_t box_aux0 = smem_u32(smem + L.aux) + static_cast<uint32_t>(ewarp * 2) * GEMM_BOX_BF16;
    const uint32_t aux_bar0 = smem_u32(&aux_bars[ewarp * 2]);
    const uint32_t row_out = static_cast<uint32_t>(lane) * (p.d_is_f32 ? 128u : 64u), row_bf16 = static_cast<uint32_t>(lane) * 64u;
    uint32_t chunk_ctr = 0;       // chunks this warp has processed: selects the staging / aux box and the aux barrier parity
    const int nch = BN / GEMM_CHUNK;
    if (lane == 0 && ewarp == 1) { if (p.has_preact) tma_prefetch_desc(&tma_x); if (has_aux) tma_prefetch_desc(&tma_aux); }
    tmem_slot_barrier();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot[0];
    const uint32_t tmem_empty_leader = CG == 2 ? mapa_u32(&tmem_empty_bar[0], leader_rank) : 0u;
    if constexpr (CS > 1) cluster_wait(); else cta_sync();
    griddep_wait();     // bias / aux reads and output writes wait for the previous kernel of the stream
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile(p, tile);
      const int m0 = tc.m_idx * (GEMM_BM * CG) + static_cast<int>(prank) * GEMM_BM;
      const int n0 = (tc.n_idx * NP + static_cast<int>(pair)) * BN;
      const int r0 = m0 + quarter * 32;
      // bias / scale of my chunks -> my strip (the previous tile's reads of the strip are complete: same warp, program order)
      __syncwarp();
      for (int j = 0, c = half; c < nch; ++j, c += 2) {
        const int col = n0 + c * GEMM_CHUNK + lane;
        const bool ok = col < p.n;
        sts_f1(sa_bias + static_cast<uint32_t>(j * GEMM_CHUNK + lane) * 4u, (p.bias != nullptr && ok) ? __ldg(p.bias + col) : 0.0f);
        if (sa_scale != 0u) sts_f1(sa_scale + static_cast<uint32_t>(j * GEMM_CHUNK + lane) * 4u, ok ? __ldg(p.scale + col) : 1.0f);
      }
      __syncwarp();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(quarter * 32) << 16);
      // One 32-column chunk per iteration: the accumulator words arrive with ONE tcgen05.ld (a TMEM load has ~300 cycles of
      // latency whatever its width: 8-column loads in a rolled loop were 2x slower), the chunk is processed by straight-line
      // code specialised at compile time, and everything the NEXT chunk needs is requested before the current one is processed:
      // its TMEM load (second register buffer) and its aux box (TMA into the warp's other aux box).  Staging boxes alternate, so
      // a chunk is staged while the previous store is still reading its box.  Iteration -1 of the CTA's first tile is a dry run on
      // zeros while the main loop is still computing: it pulls the epilogue code into the instruction cache (nothing is stored).
      constexpr bool HAS_AUX = EPI == EPI_AUX_ADD || EPI == EPI_AUX_GELUGRAD || EPI == EPI_AUX_MUL || EPI == EPI_BN_ADD_RELU || EPI == EPI_GENERIC;
      const bool use_aux = HAS_AUX && has_aux;
      const int my_chunks = (nch - half + 1) / 2;
      uint32_t r[32], rn[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) { r[i] = 0u; rn[i] = 0u; }
      const bool warm = tile == first_tile;
#pragma unroll 1
      for (int j = warm ? -1 : 0; j < my_chunks; ++j) {
        const bool live = j >= 0;
        const int jj = live ? j : 0;
        const int c = half + 2 * jj;
        const uint32_t ab = chunk_ctr & 1u;                          // aux box / barrier of this chunk
        const uint32_t sbuf = p.nbuf == 2 ? (chunk_ctr & 1u) : 0u;   // staging box of this chunk
        if (j == 0) {
          // this tile's first aux box is requested before the accumulator wait (its latency hides behind the main loop)
          if (use_aux && lane == 0) {
            mbar_arrive_expect_tx(&aux_bars[ewarp * 2 + ab], GEMM_BOX_BF16);
            tma_load_2d(smem + L.aux + (ewarp * 2 + ab) * GEMM_BOX_BF16, &tma_aux, &aux_bars[ewarp * 2 + ab], n0 + c * GEMM_CHUNK, r0);
          }
          mbar_wait(&tmem_full_bar[acc], acc_phase);
          tc_fence_after();
          if (tracer && tile == first_tile) trace_stamp(7);
          if (tracer) trace_stamp(8);     // last tile's accumulator ready
          tmem_ld_32x32(taddr + static_cast<uint32_t>(c * GEMM_CHUNK), r);
        }
        if (live) {
          if (use_aux && j + 1 < my_chunks && lane == 0) {
            // the other aux box was last read two chunks ago (same warp, program order)
            mbar_arrive_expect_tx(&aux_bars[ewarp * 2 + (ab ^ 1u)], GEMM_BOX_BF16);
            tma_load_2d(smem + L.aux + (ewarp * 2 + (ab ^ 1u)) * GEMM_BOX_BF16, &tma_aux, &aux_bars[ewarp * 2 + (ab ^ 1u)],
                        n0 + (c + 2) * GEMM_CHUNK, r0);
          }
          tmem_ld_wait();                     // covers the load issued one iteration ago (into rn; into r for j == 0)
          if (tracer && j == 0) trace_stamp(12);
          if (j > 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = rn[i];
          }
          if (j + 1 < my_chunks) {
            tmem_ld_32x32(taddr + static_cast<uint32_t>((c + 2) * GEMM_CHUNK), rn);     // arrives while this chunk is processed
          } else {
            // my share of the accumulator stage is in registers: hand the stage back to the MMA warp (of the leader) right away
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_cluster(tmem_empty_leader + static_cast<uint32_t>(acc) * 8u);
              else                   mbar_arrive(&tmem_empty_bar[acc]);
            }
          }
          // my staging box: with two boxes the store issued two chunks ago must have been read, with one the previous store
          if (lane == 0) { if (p.nbuf == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
          if (use_aux) mbar_wait_addr(aux_bar0 + ab * 8u, (chunk_ctr >> 1) & 1u);
          __syncwarp();
        }
        const uint32_t sb = sa_bias + static_cast<uint32_t>(jj * GEMM_CHUNK) * 4u;
        const uint32_t ss = sa_scale != 0u ? sa_scale + static_cast<uint32_t>(jj * GEMM_CHUNK) * 4u : 0u;
        const uint32_t orow = box_out0 + sbuf * out_box_bytes + row_out;
        const uint32_t xrow = box_x0 + sbuf * GEMM_BOX_BF16 + row_bf16;
        const uint32_t arow = box_aux0 + ab * GEMM_BOX_BF16 + row_bf16;
#pragma unroll
        for (int g = 0; g < 4; ++g) epilogue_octet<EPI>(p, &r[8 * g], g, sb, ss, arow, xrow, orow, swz64, swz128);
        if (tracer && live && j == 0) trace_stamp(13);
        fence_proxy_async_smem();           // make the staged box visible to the TMA engine
        __syncwarp();
        const int c0 = n0 + c * GEMM_CHUNK;
        if (live && lane == 0) {
          if (c0 < p.n && r0 < p.m) {
            if (EPI == EPI_F32 || (EPI == EPI_GENERIC && p.d_is_f32)) {
              if (p.reduce_add) tma_reduce_add_2d_hint(&tma_d, smem + L.out + (ewarp * p.nbuf + sbuf) * GEMM_BOX_F32, c0, r0, p.d_policy ? p.d_policy : L2_EVICT_NORMAL);
              else              tma_store_2d_hint(&tma_d, smem + L.out + (ewarp * p.nbuf + sbuf) * GEMM_BOX_F32, c0, r0, p.d_policy ? p.d_policy : L2_EVICT_NORMAL);
            } else {
              tma_store_2d(&tma_d, smem + L.out + (ewarp * p.nbuf + sbuf) * GEMM_BOX_BF16, c0, r0);
              if (EPI == EPI_GELU_PRE || EPI == EPI_GELU_GRADPRE || (EPI == EPI_GENERIC && p.has_preact))
                tma_store_2d(&tma_x, smem + L.x + (ewarp * p.nbuf + sbuf) * GEMM_BOX_BF16, c0, r0);
            }
          }
          tma_store_commit();               // one (possibly empty) group per chunk: wait_group.read 1 counts chunks
          if (tracer && j == 0) trace_stamp(16);
        }
        if (live) ++chunk_ctr;
      }
      if (++acc == GEMM_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
    if (tracer) trace_stamp(9);
    if (lane == 0) tma_store_wait_read();   // smem may be released; the writes themselves complete before the grid does
    if (tracer) trace_stamp(10);
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncwarp();
  if constexpr (CS > 1) {
    // the other CTAs' smem / barriers / TMEM stay alive until every thread of the cluster is done with them
    cluster_arrive_relaxed();
    cluster_wait();
  } else {
    cta_sync();
  }
  if (warp == 1) {
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot[0];
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    else                   tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
#ifdef VB_GEMM_TRACE
  if (threadIdx.x == 0) trace_stamp(11);
  if (threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); s_trace[23] = static_cast<long long>(g); }
  __syncthreads();
  if (p.trace != nullptr && threadIdx.x < 24) p.trace[blockIdx.x * 24 + threadIdx.x] = s_trace[threadIdx.x];
#endif
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static long long* g_trace = nullptr;

static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    g_num_sms = n;
  }
  return g_num_sms;
}

// experiment knobs: read once from the environment, overridable at run time through vb_gemm_set_knob (tools/gemm_mainloop.py)
struct GemmKnobs { int occ1, max_bn, np, cg, debug_mode, stages, no_l2_hints, compact, pdl_late, smem_kb, max_stages, nbuf; };
static int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
static GemmKnobs& knobs() {
  static GemmKnobs k = {env_int("VB_GEMM_OCC1", 0), env_int("VB_GEMM_MAX_BN", 256), env_int("VB_GEMM_NP", 0), env_int("VB_GEMM_CG", 0),
                        env_int("VB_GEMM_DEBUG", 0), env_int("VB_GEMM_STAGES", 0), env_int("VB_GEMM_NO_L2_HINTS", 0),
                        env_int("VB_GEMM_COMPACT", 0), env_int("VB_GEMM_PDL_LATE", 1), env_int("VB_GEMM_SMEM_KB", 0),
                        env_int("VB_GEMM_MAX_STAGES", 10), env_int("VB_GEMM_NBUF", 2)};
  return k;
}
// Residency target per SM.  VB_GEMM_COMPACT (knob "compact"): 0 = one CTA per SM for every GEMM (default: deep operand ring,
// double-buffered epilogue, up to 200 registers), 1 = tiles up to 128 wide use the compact configuration (<= 112 KB, <= 96
// registers) so that CTAs of two kernels can share an SM, 2 = compact for fp32 outputs (weight gradients) only, 3 = compact for
// bf16 outputs only.
static int gemm_occupancy(const vb_gemm_args& a, int bn) {
  const int mode = knobs().occ1 ? 0 : knobs().compact;
  if (bn > 128 || mode == 0 || a.conv_kh > 0) return 1;
  if (mode == 2) return a.d_is_f32 ? 2 : 1;
  if (mode == 3) return a.d_is_f32 ? 1 : 2;
  return 2;
}

// Co-resident clusters of `cs` CTAs (one CTA per SM).  Clusters cannot straddle a GPC and the B200's GPCs do not all hold a
// multiple of four SMs, so fewer than 148 / cs clusters of four fit; asked of the driver once per cluster size.
static int max_clusters(int cs, int occupancy) {
  static int cached[5][3] = {};
  if (cs <= 1) return num_sms() * occupancy;
  if (cached[cs][occupancy] == 0) {
    int n = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(num_sms() / cs * cs));
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = gemm_smem_limit(occupancy);
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = static_cast<unsigned>(cs);
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (occupancy == 2) {
      if (cs == 2) {
        auto kern = gemm_bf16_kernel<false, false, 2, 1, 2, EPI_GENERIC>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_limit(2));
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      } else {
        auto kern = gemm_bf16_kernel<false, false, 2, 2, 2, EPI_GENERIC>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_limit(2));
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      }
    } else {
      if (cs == 2) {
        auto kern = gemm_bf16_kernel<false, false, 2, 1, 1, EPI_GENERIC>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_limit(1));
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      } else {
        auto kern = gemm_bf16_kernel<false, false, 2, 2, 1, EPI_GENERIC>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_limit(1));
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      }
    }
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms() / cs * 7 / 8 * occupancy; }
    cached[cs][occupancy] = n;
  }
  return cached[cs][occupancy];
}

static int gemm_nbuf(const vb_gemm_args& a, int bn) { return (gemm_occupancy(a, bn) == 1 && knobs().nbuf == 2) ? 2 : 1; }
static int gemm_stages(const vb_gemm_args& a, int bn, int cg) {
  const int occ = gemm_occupancy(a, bn);
  const int bnl = bn / cg;
  const GemmSmem z = gemm_smem(0, bnl, bn, a.d_is_f32 != 0, a.d_preact != nullptr, a.scale != nullptr, a.aux_mode != VB_AUX_NONE, gemm_nbuf(a, bn));
  // VB_GEMM_SMEM_KB: leave shared memory on the SM to co-resident blocks of other streams' kernels (LayerNorm backward, ...)
  int limit = gemm_smem_limit(occ);
  if (occ == 1 && knobs().smem_kb > 0 && knobs().smem_kb * 1024 < limit) limit = knobs().smem_kb * 1024;
  int stages = (limit - 1024 /*alignment slack*/ - static_cast<int>(z.total)) / (GEMM_A_BYTES + bnl * 128);
  // A kernel timed alone runs its main loop at the same rate with 3 and with 8 stages (tools/gemm_mainloop.py), but inside the
  // training step, where four streams share L2 and HBM, the depth absorbs the latency jitter: 4 stages 5.08 ms/step, 6 -> 4.76,
  // 8 -> 4.73, 10 -> 4.70 (VB_GEMM_MAX_STAGES)
  int cap = knobs().max_stages;
  if (cap > GEMM_MAX_STAGES) cap = GEMM_MAX_STAGES;
  if (stages > cap) stages = cap;
  return stages;
}

template <bool A_MN, bool B_MN, int CG, int NP, int OCC, int EPI, bool CONV = false>
static int launch_gemm(const vb_gemm_args& a, int bn, int splits, cudaStream_t stream) {
  static_assert(!CONV || (!A_MN && !B_MN), "implicit-GEMM convolutions read K-major operands");
  constexpr int CS = CG * NP;
  const int bnl = bn / CG;
  CUtensorMap map_a, map_b, map_d, map_x, map_aux;
  int rc;
  // K-major operand: global [rows, K] -> box {64 (k), rows_per_cta}; MN-major: global [K, rows] -> box {64 (mn), 64 (k)}
  // (Tried: describing an MN-major operand as a 3-D tensor (64 | k | piece) so that ONE instruction fetches every 64-wide piece
  // of a tile.  Measured slower on B200 -- weight gradients 1 031 -> 1 270 us per step, profiles/r02_gemm_experiments.md (gpurun_out/r02v_*) -- and
  // removed: one 2-D load per piece.)
  if (CONV)      rc = make_tensor_map_im2col(&map_a, a.a, a.conv_n, a.conv_h, a.conv_w, a.conv_c, a.conv_kh, a.conv_kw, a.conv_stride,
                                             a.conv_pad, GEMM_BM / NP);
  else if (A_MN) rc = make_tensor_map_2d(&map_a, a.a, /*inner*/ a.m, /*outer*/ a.k, a.lda, 64, GEMM_BK);
  else           rc = make_tensor_map_2d(&map_a, a.a, a.k, a.m, a.lda, GEMM_BK, GEMM_BM / NP);   // NP = 2: multicast halves
  if (rc != VB_OK) return rc;
  if (B_MN) rc = make_tensor_map_2d(&map_b, a.b, a.n, a.k, a.ldb, 64, GEMM_BK);
  else      rc = make_tensor_map_2d(&map_b, a.b, a.k, a.n, a.ldb, GEMM_BK, bnl);
  if (rc != VB_OK) return rc;
  // epilogue boxes: 32 rows x 32 columns per warp (fp32: 128-byte rows, 128B swizzle; bf16: 64-byte rows, 64B swizzle)
  if (a.d_is_f32) rc = make_tensor_map_2d_f32(&map_d, a.d, a.n, a.m, a.ldd, GEMM_CHUNK, 32);
  else            rc = make_tensor_map_2d_sw64(&map_d, a.d, a.n, a.m, a.ldd, GEMM_CHUNK, 32);
  if (rc != VB_OK) return rc;
  if (a.d_preact != nullptr) {
    rc = make_tensor_map_2d_sw64(&map_x, a.d_preact, a.n, a.m, a.ld_preact, GEMM_CHUNK, 32);
    if (rc != VB_OK) return rc;
  } else {
    map_x = map_d;
  }
  if (a.aux_mode != VB_AUX_NONE) {
    rc = make_tensor_map_2d_sw64(&map_aux, a.aux, a.n, a.m, a.ld_aux, GEMM_CHUNK, 32);
    if (rc != VB_OK) return rc;
  } else {
    map_aux = map_d;
  }

  GemmKernelParams p;
  p.conv_hw_out = p.conv_w_out = p.conv_cblocks = p.conv_kw = 1; p.conv_stride = 1; p.conv_pad = 0;
  if (CONV) {
    const int ho = (a.conv_h + 2 * a.conv_pad - a.conv_kh) / a.conv_stride + 1, wo = (a.conv_w + 2 * a.conv_pad - a.conv_kw) / a.conv_stride + 1;
    p.conv_hw_out = ho * wo; p.conv_w_out = wo; p.conv_stride = a.conv_stride; p.conv_pad = a.conv_pad; p.conv_kw = a.conv_kw;
    p.conv_cblocks = a.conv_c / GEMM_BK;
  }
  p.scale = a.scale; p.bias = a.bias;
  p.aux = static_cast<const __nv_bfloat16*>(a.aux); p.ld_aux = a.ld_aux;
  p.m = a.m; p.n = a.n; p.k = a.k;
  p.bn = bn;
  p.d_is_f32 = a.d_is_f32; p.act = a.act; p.aux_mode = a.aux_mode;
  p.has_preact = a.d_preact != nullptr;
  p.preact_grad = a.preact_grad;
  p.m_tiles = (a.m + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  p.n_tiles = ((a.n + bn - 1) / bn + NP - 1) / NP;   // N steps of a whole cluster (NP tiles side by side)
  const int total_kb = (a.k + GEMM_BK - 1) / GEMM_BK;
  if (splits > total_kb) splits = total_kb;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  p.splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  p.reduce_add = (a.accumulate || p.splits > 1) ? 1 : 0;
  {
    const uint64_t mn = static_cast<uint64_t>(p.m_tiles) * p.n_tiles, tiles = mn * p.splits;
    if (tiles * mn >= (1ull << 32)) {
      vb_set_last_error("vb_gemm_bf16", "problem too large for the 32-bit tile decoder");
      return VB_ERR_UNSUPPORTED;
    }
    p.magic_m = p.m_tiles == 1 ? 0u : static_cast<uint32_t>(((1ull << 32) + p.m_tiles - 1) / p.m_tiles);
    p.magic_mn = mn == 1 ? 0u : static_cast<uint32_t>(((1ull << 32) + mn - 1) / mn);
  }
  int stages = gemm_stages(a, bn, CG);
  p.nbuf = gemm_nbuf(a, bn);
  p.pdl_late = knobs().pdl_late;
  if (stages < 2) {
    vb_set_last_error("vb_gemm_bf16", "tile configuration does not fit shared memory");
    return VB_ERR_UNSUPPORTED;
  }
  const int debug_stages = knobs().stages;
  p.debug_mode = knobs().debug_mode;
  p.trace = g_trace;
  const bool hints = knobs().no_l2_hints == 0;
  p.b_policy = (hints && a.b_streamed) ? L2_EVICT_FIRST : 0ull;     // 0 = the plain (un-hinted) instruction
  p.d_policy = (hints && a.d_streamed && a.d_is_f32) ? L2_EVICT_FIRST : 0ull;
  if (debug_stages >= 2 && debug_stages < stages) stages = debug_stages;
  p.stages = stages;
  p.tmem_cols = 32;
  while (p.tmem_cols < GEMM_ACC_STAGES * bn) p.tmem_cols *= 2;
  const GemmSmem L = gemm_smem(stages, bnl, bn, a.d_is_f32 != 0, p.has_preact != 0, a.scale != nullptr, a.aux_mode != VB_AUX_NONE, p.nbuf);
  const int smem_bytes = 1024 + static_cast<int>(L.total);

  static bool attr_set = false;   // one per instantiation
  auto kern = gemm_bf16_kernel<A_MN, B_MN, CG, NP, OCC, EPI, CONV>;
  if (!attr_set) {
    VB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_limit(OCC)));
    attr_set = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.splits;
  int groups = max_clusters(CS, OCC);
  if (a.max_ctas > 0 && a.max_ctas / CS < groups) groups = a.max_ctas / CS > 0 ? a.max_ctas / CS : 1;
  if (tiles < groups) groups = tiles;
  VB_CUDA_CHECK(launch_ex(kern, dim3(groups * CS), dim3(GEMM_THREADS), smem_bytes, stream, CS, /*pdl=*/true, map_a, map_b,
                          map_d, map_x, map_aux, p));
  return VB_OK;
}

// Tile-shape heuristic.  A CTA's time = prologue + (tiles per CTA) x k-blocks x max(tensor-pipe floor, what the SM can
// ingest, its share of what L2 can deliver) + the LAST tile's epilogue (earlier ones overlap the next main loop); the cheapest
// candidate wins.  Constants measured with tools/gemm_trace.py on B200: L2 delivers ~6300 B/clk to all SMs together, one SM
// takes ~50 (3-stage ring) to ~57 B/clk, a UMMA 128 x N x 16 (per SM) costs N/2 cycles.
struct TileChoice { int bn, cg, np, splits; };

static double tile_cost(const vb_gemm_args& a, int bn, int cg, int np, int kb, double ctas, double waves, bool split) {
  const bool f32 = a.d_is_f32 != 0;
  const int stages = gemm_stages(a, bn, cg);
  const double sms = ctas < num_sms() ? ctas : num_sms();
  const double share = ctas / sms;     // two small CTAs co-resident share the SM's ingest bandwidth and tensor pipe
  const double a_bytes = GEMM_BM * 128.0, b_bytes = (bn / cg) * 128.0;
  const double rate = stages >= 5 ? 57.0 : (stages == 4 ? 54.0 : (stages == 3 ? 50.0 : 40.0));
  const double ingest = (a_bytes + b_bytes) / rate * share;
  const double l2 = (a_bytes / np + b_bytes) * ctas / 6300.0;
  const double mma = (2.0 * bn + 16.0) * share;
  double kb_cost = ingest > mma ? ingest : mma;
  if (l2 > kb_cost) kb_cost = l2;
  double epi = 500.0 + (f32 ? 7.0 : 4.0) * bn + (split ? 2.0 * bn : 0.0);
  if (a.act == VB_ACT_GELU) epi += 3.0 * bn;
  if (a.d_preact != nullptr) epi += 3.0 * bn;
  if (a.aux_mode == VB_AUX_MUL_GELU_GRAD) epi += 5.0 * bn;
  // an epilogue longer than the next tile's main loop is exposed on every tile, not just the last
  const double main_loop = kb * kb_cost;
  const double per_tile = main_loop > epi ? main_loop : epi;
  return 900.0 + (waves - 1.0) * per_tile + main_loop + epi;
}

static bool tile_legal(const vb_gemm_args& a, int bn, int cg) {
  if (bn > knobs().max_bn) return false;   // experiments: compact tiles only
  if (cg == 1) return bn == 64 || bn == 128;
  if (a.b_mn_major) return bn == 128 || bn == 256;              // 64-wide MN pieces per CTA
  if (a.a_mn_major) return bn == 128;                           // (MN, K): API completeness only
  return bn % 32 == 0 && bn >= 64 && bn <= 256;
}

static TileChoice pick_config(const vb_gemm_args& a) {
  const int force_np = knobs().np, force_cg = knobs().cg;
  const int total_kb = (a.k + GEMM_BK - 1) / GEMM_BK;
  const bool can_split = a.d_is_f32 && a.accumulate;
  double best_cost = 1e30;
  TileChoice best = {128, a.m > GEMM_BM ? 2 : 1, 1, 1};
  const int bns[7] = {128, 96, 160, 192, 224, 256, 64};
  for (int pass = 0; pass < 2 && best_cost > 1e29; ++pass) {     // pass 1: the block_n hint was not legal, ignore it
    for (int cg = 2; cg >= 1; --cg) {
      if (cg == 2 && a.m <= GEMM_BM) continue;                   // a pair would leave its second CTA without rows
      if (force_cg != 0 && a.m > GEMM_BM && cg != force_cg) continue;
      const int m_tiles = (a.m + GEMM_BM * cg - 1) / (GEMM_BM * cg);
      for (int np = 1; np <= cg; ++np) {
        if (force_np != 0 && cg == 2 && np != force_np) continue;
        for (int bi = 0; bi < 7; ++bi) {
          const int bn = bns[bi];
          if (!tile_legal(a, bn, cg)) continue;
          int clusters = max_clusters(cg * np, gemm_occupancy(a, bn));
          if (a.max_ctas > 0 && a.max_ctas / (cg * np) < clusters) clusters = a.max_ctas / (cg * np) > 0 ? a.max_ctas / (cg * np) : 1;
          if (pass == 0 && a.block_n != 0 && a.block_n != bn) continue;
          if (bn > 64 && a.n <= bn / 2 && pass == 0 && a.block_n == 0) continue;  // do not waste most of a tile
          const int n_tiles = ((a.n + bn - 1) / bn + np - 1) / np;                 // cluster steps along N
          if (np == 2 && (a.n + bn - 1) / bn < 2) continue;
          const int max_s = can_split ? 16 : 1;
          for (int s = 1; s <= max_s; s *= 2) {
            if (a.splits != 0 && a.splits != s) continue;
            if (s > 1 && total_kb / s < 4) break;
            const long tiles = static_cast<long>(m_tiles) * n_tiles * s;
            const long waves = (tiles + clusters - 1) / clusters;
            const int kb = (total_kb + s - 1) / s;
            const double ctas = static_cast<double>(tiles < clusters ? tiles : clusters) * cg * np;
            const double cost = tile_cost(a, bn, cg, np, kb, ctas, static_cast<double>(waves), s > 1);
            if (cost < best_cost * 0.97) { best_cost = cost; best = {bn, cg, np, s}; }
          }
        }
      }
    }
  }
  return best;
}

// Which compiled epilogue serves this call (anything unusual -> the generic kernel with run-time flags)
static int pick_epilogue(const vb_gemm_args& a) {
  static const bool generic_only = env_int("VB_GEMM_GENERIC_EPI", 0) != 0;
  if (generic_only) return EPI_GENERIC;
  if (a.scale != nullptr) {
    const bool bn = a.bias != nullptr && !a.d_is_f32 && a.d_preact == nullptr && a.act == VB_ACT_RELU;
    if (bn && a.aux_mode == VB_AUX_NONE) return EPI_BN_RELU;
    if (bn && a.aux_mode == VB_AUX_ADD) return EPI_BN_ADD_RELU;
    return EPI_GENERIC;
  }
  if (a.d_is_f32) return (a.bias == nullptr && a.act == VB_ACT_NONE && a.aux_mode == VB_AUX_NONE) ? EPI_F32 : EPI_GENERIC;
  if (a.d_preact != nullptr)
    return (a.act == VB_ACT_GELU && a.aux_mode == VB_AUX_NONE) ? (a.preact_grad ? EPI_GELU_GRADPRE : EPI_GELU_PRE) : EPI_GENERIC;
  if (a.aux_mode != VB_AUX_NONE) {
    if (a.bias != nullptr || a.act != VB_ACT_NONE) return EPI_GENERIC;
    return a.aux_mode == VB_AUX_ADD ? EPI_AUX_ADD : (a.aux_mode == VB_AUX_MUL ? EPI_AUX_MUL : EPI_AUX_GELUGRAD);
  }
  return a.act == VB_ACT_NONE ? EPI_BIAS : EPI_GENERIC;
}

// Instantiated combinations: the specialised epilogues only for the operand layouts the training step uses them with
// (forward K-major/K-major, dgrad K-major/MN-major, wgrad MN-major/MN-major) and only for CTA pairs; everything else is generic.
template <int CG, int NP, int OCC>
static int dispatch_major(const vb_gemm_args& a, int bn, int splits, cudaStream_t s) {
  const int epi = CG == 2 ? pick_epilogue(a) : EPI_GENERIC;
  if (a.conv_kh > 0) {       // implicit-GEMM convolution (one CTA per SM): BatchNorm + ReLU epilogue compiled in, anything else generic
    if constexpr (OCC == 1) {
      if constexpr (CG == 2) {
        if (epi == EPI_BN_RELU) return launch_gemm<false, false, CG, NP, 1, EPI_BN_RELU, true>(a, bn, splits, s);
      }
      return launch_gemm<false, false, CG, NP, 1, EPI_GENERIC, true>(a, bn, splits, s);
    } else { vb_set_last_error("vb_gemm_bf16", "convolution kernels are built for one CTA per SM"); return VB_ERR_UNSUPPORTED; }
  }
  if constexpr (CG == 2) {
    if (!a.a_mn_major && !a.b_mn_major) {
      if (epi == EPI_BIAS) return launch_gemm<false, false, CG, NP, OCC, EPI_BIAS>(a, bn, splits, s);
      if constexpr (OCC == 1) {      // the 1x1 convolutions of the ResNet trunk (the activation itself is the operand)
        if (epi == EPI_BN_RELU) return launch_gemm<false, false, CG, NP, OCC, EPI_BN_RELU>(a, bn, splits, s);
        if (epi == EPI_BN_ADD_RELU) return launch_gemm<false, false, CG, NP, OCC, EPI_BN_ADD_RELU>(a, bn, splits, s);
      }
      if (epi == EPI_GELU_PRE) return launch_gemm<false, false, CG, NP, OCC, EPI_GELU_PRE>(a, bn, splits, s);
      if (epi == EPI_GELU_GRADPRE) return launch_gemm<false, false, CG, NP, OCC, EPI_GELU_GRADPRE>(a, bn, splits, s);
    } else if (!a.a_mn_major && a.b_mn_major) {
      if (epi == EPI_BIAS) return launch_gemm<false, true, CG, NP, OCC, EPI_BIAS>(a, bn, splits, s);
      if (epi == EPI_AUX_ADD) return launch_gemm<false, true, CG, NP, OCC, EPI_AUX_ADD>(a, bn, splits, s);
      if (epi == EPI_AUX_GELUGRAD) return launch_gemm<false, true, CG, NP, OCC, EPI_AUX_GELUGRAD>(a, bn, splits, s);
      if (epi == EPI_AUX_MUL) return launch_gemm<false, true, CG, NP, OCC, EPI_AUX_MUL>(a, bn, splits, s);
    } else if (a.a_mn_major && a.b_mn_major) {
      if (epi == EPI_F32) return launch_gemm<true, true, CG, NP, OCC, EPI_F32>(a, bn, splits, s);
      if (epi == EPI_BIAS) return launch_gemm<true, true, CG, NP, OCC, EPI_BIAS>(a, bn, splits, s);   // bf16 weight gradients (switch exchange)
    }
  }
  if (a.a_mn_major && a.b_mn_major) return launch_gemm<true, true, CG, NP, OCC, EPI_GENERIC>(a, bn, splits, s);
  if (a.a_mn_major) return launch_gemm<true, false, CG, NP, OCC, EPI_GENERIC>(a, bn, splits, s);
  if (a.b_mn_major) return launch_gemm<false, true, CG, NP, OCC, EPI_GENERIC>(a, bn, splits, s);
  return launch_gemm<false, false, CG, NP, OCC, EPI_GENERIC>(a, bn, splits, s);
}

template <int CG, int NP>
static int dispatch_occ(const vb_gemm_args& a, const TileChoice& c, cudaStream_t s) {
  if (gemm_occupancy(a, c.bn) == 2) return dispatch_major<CG, NP, 2>(a, c.bn, c.splits, s);
  return dispatch_major<CG, NP, 1>(a, c.bn, c.splits, s);
}

}  // namespace vb

// diagnostic: resident blocks per SM / co-resident 2-CTA clusters the runtime reports for the compact pair kernel at a given
// dynamic shared-memory size (tools/gemm_occupancy.py)
extern "C" int vb_gemm_debug_occupancy(int smem_bytes, int* blocks_per_sm, int* clusters) {
  auto kern = vb::gemm_bf16_kernel<false, false, 2, 1, 2, vb::EPI_GENERIC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, vb::GEMM_THREADS, smem_bytes) != cudaSuccess) return VB_ERR_CUDA;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(vb::GEMM_THREADS); cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  if (cudaOccupancyMaxActiveClusters(clusters, kern, &cfg) != cudaSuccess) return VB_ERR_CUDA;
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  fprintf(stderr, "regs %d static smem %zu maxDyn %d\n", fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes);
  return VB_OK;
}

extern "C" int vb_gemm_set_knob(const char* name, int value) {
  using namespace vb;
  GemmKnobs& k = knobs();
  const struct { const char* n; int* v; } tab[] = {{"occ1", &k.occ1}, {"max_bn", &k.max_bn}, {"np", &k.np}, {"cg", &k.cg},
                                                   {"debug_mode", &k.debug_mode}, {"stages", &k.stages}, {"no_l2_hints", &k.no_l2_hints},
                                                   {"compact", &k.compact}, {"pdl_late", &k.pdl_late}, {"smem_kb", &k.smem_kb},
                                                   {"max_stages", &k.max_stages}, {"nbuf", &k.nbuf}};
  for (const auto& e : tab)
    if (strcmp(e.n, name) == 0) { *e.v = value; return VB_OK; }
  vb_set_last_error("vb_gemm_set_knob", "unknown knob");
  return VB_ERR_BAD_ARG;
}

extern "C" int vb_gemm_set_trace(void* device_buffer) {
  vb::g_trace = static_cast<long long*>(device_buffer);
  return VB_OK;
}

extern "C" int vb_gemm_bf16(const vb_gemm_args* args, void* stream) {
  using namespace vb;
  VB_REQUIRE(args != nullptr, "null args");
  const vb_gemm_args& a = *args;
  VB_REQUIRE(a.a && a.b && a.d, "a, b and d must be non-null device pointers");
  VB_REQUIRE(a.m > 0 && a.n > 0 && a.k > 0, "m, n, k must be positive");
  VB_REQUIRE(a.n % 8 == 0, "n must be a multiple of 8");
  if (a.conv_kh > 0) {
    VB_REQUIRE(a.conv_kw > 0 && a.conv_stride > 0 && a.conv_pad >= 0 && a.conv_n > 0 && a.conv_h > 0 && a.conv_w > 0 && a.conv_c > 0,
               "convolution geometry must be positive");
    VB_REQUIRE(a.conv_c % 64 == 0, "implicit-GEMM convolution needs a multiple of 64 input channels");
    VB_REQUIRE(!a.a_mn_major && !a.b_mn_major, "implicit-GEMM convolution takes an NHWC activation and a [Cout, kh*kw*Cin] weight");
    VB_REQUIRE(a.conv_kh <= 16 && a.conv_kw <= 16 && a.conv_pad < 128 && a.conv_stride <= 8, "window / padding / stride out of the TMA im2col range");
    const long long ho = (a.conv_h + 2 * a.conv_pad - a.conv_kh) / a.conv_stride + 1, wo = (a.conv_w + 2 * a.conv_pad - a.conv_kw) / a.conv_stride + 1;
    VB_REQUIRE(ho > 0 && wo > 0 && a.m == (long long)a.conv_n * ho * wo && a.k == a.conv_kh * a.conv_kw * a.conv_c,
               "m must be n*ho*wo and k must be kh*kw*c of the convolution");
    VB_REQUIRE(a.splits <= 1 && !a.accumulate, "no split-K for convolutions");
  }
  VB_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "lda/ldb must be multiples of 8 elements (16-byte TMA strides)");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(a.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.b) & 15) == 0,
             "a/b must be 16-byte aligned");
  VB_REQUIRE(a.ldd % (a.d_is_f32 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(a.d) & 15) == 0, "d must be 16-byte aligned rows");
  VB_REQUIRE(a.d_preact == nullptr || (a.ld_preact % 8 == 0 && (reinterpret_cast<uintptr_t>(a.d_preact) & 15) == 0), "d_preact alignment");
  VB_REQUIRE(a.aux_mode == VB_AUX_NONE || (a.aux != nullptr && a.ld_aux % 8 == 0 && (reinterpret_cast<uintptr_t>(a.aux) & 15) == 0), "aux missing or misaligned");
  VB_REQUIRE(a.scale == nullptr || (reinterpret_cast<uintptr_t>(a.scale) & 15) == 0, "scale alignment");
  VB_REQUIRE(a.bias == nullptr || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "bias alignment");
  VB_REQUIRE(a.block_n == 0 || (a.block_n % 32 == 0 && a.block_n >= 64 && a.block_n <= 256), "block_n must be 0 or a multiple of 32 in [64, 256]");
  VB_REQUIRE(a.splits >= 0 && (a.splits <= 1 || (a.d_is_f32 && a.accumulate)), "split-K needs an fp32 accumulating output");
  VB_REQUIRE(!(a.accumulate && !a.d_is_f32), "accumulate needs an fp32 output");
  VB_REQUIRE(!(a.d_is_f32 && a.d_preact != nullptr), "d_preact only with a bf16 output");
  VB_REQUIRE(!a.preact_grad || (a.d_preact != nullptr && a.act == VB_ACT_GELU), "preact_grad stores GELU'(pre-activation): needs d_preact and act = GELU");
  VB_REQUIRE(a.aux_mode >= VB_AUX_NONE && a.aux_mode <= VB_AUX_MUL, "unknown aux_mode");
  VB_REQUIRE(!(a.d_is_f32 && a.aux_mode != VB_AUX_NONE), "aux only with a bf16 output");
  // VB_GEMM_MAX_CTAS: default cap of the persistent grid (data-parallel runs leave a few SMs to the NCCL kernels, so that a
  // GEMM sized for the whole chip does not have to wait for SMs a collective is sitting on)
  static const int env_max_ctas = getenv("VB_GEMM_MAX_CTAS") ? atoi(getenv("VB_GEMM_MAX_CTAS")) : 0;
  vb_gemm_args capped;
  if (a.max_ctas == 0 && env_max_ctas > 0) {
    capped = a;
    capped.max_ctas = env_max_ctas;
    return vb_gemm_bf16(&capped, stream);
  }
  const TileChoice c = pick_config(a);
  static const bool verbose = getenv("VB_GEMM_VERBOSE") != nullptr;
  if (verbose)
    fprintf(stderr, "vb_gemm %dx%dx%d a_mn=%d b_mn=%d f32=%d -> bn=%d cg=%d np=%d splits=%d stages=%d (clusters %d)\n", a.m, a.n, a.k,
            a.a_mn_major, a.b_mn_major, a.d_is_f32, c.bn, c.cg, c.np, c.splits,
            gemm_stages(a, c.bn, c.cg), max_clusters(c.cg * c.np, gemm_occupancy(a, c.bn)));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (c.cg == 1) return dispatch_occ<1, 1>(a, c, s);
  if (c.np == 2) return dispatch_occ<2, 2>(a, c, s);
  return dispatch_occ<2, 1>(a, c, s);
}
