// tcgen05 / TMEM GEMM for sm_100a, fed by TMA.  One persistent, warp-specialised kernel:
//
//   warp 0      TMA producer      global -> 128B-swizzled smem ring (mbarrier full/empty)
//   warp 1      MMA issuer        one thread issues tcgen05.mma (bf16 -> fp32 in TMEM)
//   warps 2..9  epilogue          tcgen05.ld accumulator -> scale/bias/aux/activation -> swizzled smem -> TMA store
//
// CG = 2 (the normal case): the kernel runs as CTA PAIRS (2-wide clusters = the two SMs of a TPC) and issues
// tcgen05.mma.cta_group::2 with M = 256: each CTA holds its own 128 rows of A, HALF of the B tile and its own 128
// accumulator rows.  At the ViLBERT shapes (M = 1600 / 2048) the main loop is bound by the bytes an SM can pull from L2
// (~50-60 B/clk/SM measured, 64 nominal), not by the tensor pipe, so halving the B bytes per SM is what moves the needle:
// a 256 x BN pair tile ingests (16 KB + BN*64 B) per SM per 64-deep k-block, against (16 KB + BN*128 B) for a lone CTA
// with the same output per SM.  The tile width BN (64 / 96 / 128 / 192 / 256) is picked per problem so that the pair
// tiles fill the 74 TPCs in as few waves as possible.  CG = 1 (single CTA, M = 128) remains for one-row-block problems
// (poolers, classifier).
//
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.  Operands may be K-major or
// MN-major (descriptor + TMA box change only), so forward, dgrad and wgrad of nn.Linear (reference
// models/vilbert_facebook_arch.py:127-129 etc.) all run here without a transposed copy of anything.  The kernel is
// launched with programmatic dependent launch: its prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps
// the tail of the previous kernel on the stream, and griddepcontrol.wait precedes the first global access.
// See include/vilbert_b200.h for the ABI.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "launch.h"
#include "../../include/vilbert_b200.h"
#include "tensormap.h"

namespace vb {

constexpr int GEMM_BM = 128;        // output rows owned by one CTA (a pair covers 256)
constexpr int GEMM_BK = 64;         // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_EPI_WARPS = 8;   // two per TMEM sub-partition; they alternate over the 32-column chunks of a panel
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_ACC_STAGES = 2;
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_CHUNK = 32;            // epilogue granule: 32 accumulator columns = one staging box
// Tiles up to 128 columns wide use a compact configuration (<= 112 KB smem, <= 102 registers, 256 TMEM columns, two-chunk
// staging panels) meant to let TWO CTAs share an SM, so that kernels of the concurrent text / visual / weight-gradient
// streams overlap each other's prologue and epilogue.  Measured (tools/gemm_occupancy.py): the runtime still reports ONE
// resident block per SM for this kernel whatever its shared-memory, register or thread budget (also with 6 warps / 96
// registers), so the co-residency does not materialise on this driver; the compact configuration is kept because it is
// the faster one in the full step (5.95 vs 6.15 ms).  Wider tiles need all 512 TMEM columns anyway.
#ifdef VB_GEMM_OCC1
__host__ __device__ constexpr int gemm_occupancy(int bn) { return 1; }
#else
__host__ __device__ constexpr int gemm_occupancy(int bn) { return bn <= 128 ? 2 : 1; }
#endif
__host__ __device__ constexpr int gemm_panel_chunks(int bn) { return gemm_occupancy(bn) == 2 ? 2 : 4; }   // chunks staged (and stored) together
constexpr int GEMM_BOX_BF16 = GEMM_BM * GEMM_CHUNK * 2;   // 128 rows x 64 B, 64B-swizzled
constexpr int GEMM_BOX_F32 = GEMM_BM * GEMM_CHUNK * 4;    // 128 rows x 128 B, 128B-swizzled
#ifdef VB_GEMM_TRACE
constexpr int GEMM_SMEM_SLACK = 256;   // room for the static trace buffer
#else
constexpr int GEMM_SMEM_SLACK = 0;
#endif
__host__ __device__ constexpr int gemm_smem_limit(int bn) { return (gemm_occupancy(bn) == 2 ? 112 * 1024 : 227 * 1024) - GEMM_SMEM_SLACK; }

struct GemmKernelParams {
  const float* scale;
  const float* bias;
  int m, n, k;
  int d_is_f32, reduce_add, act, aux_mode, has_preact;
  int splits, kb_per_split;
  int m_tiles, n_tiles;
  int stages;       // depth of the operand ring
  int out_bytes;    // staging bytes for one output panel
  int x_bytes;      // staging bytes for the aux-in / preact-out panel (0 = unused)
  int debug_mode;   // profiling only (VB_GEMM_DEBUG): 1 = no MMA issue, 2 = no TMA loads; results are garbage
  uint32_t magic_m, magic_mn;   // fast_div multipliers for m_tiles and m_tiles * n_tiles
  unsigned long long b_policy, d_policy;   // L2 eviction hints for the B loads / D stores
  long long* trace; // profiling only (vb_gemm_set_trace): 16 clock64 stamps per CTA, NULL in production
};

// compiled in only with -DVB_GEMM_TRACE (tools/gemm_trace.py builds its own copy of the library): even a predicated-off
// stamp is instructions on the cold, instruction-fetch-bound path of a 5 us kernel
#ifdef VB_GEMM_TRACE
__shared__ long long s_trace[24];   // stamps go to shared memory and are dumped at exit: global stores would perturb
#endif
__device__ __forceinline__ void trace_stamp(const GemmKernelParams& p, int slot) {
#ifdef VB_GEMM_TRACE
  s_trace[slot] = clock64();
#endif
}

// Kernel parameters live in the constant bank, and ptxas re-loads them at every use because such loads are "free" -- but a
// cold LDC / LDCU costs ~100 cycles, and a 5 us kernel whose epilogue tests four flags per 16 columns pays that latency in
// a dependent chain, dozens of times (measured: ~1300 of ~2000 epilogue cycles).  pin() forces a value into a register
// once, at kernel entry, where all the loads overlap each other and the TMEM allocation.
__device__ __forceinline__ void pin(int& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void pin(uint32_t& x) { asm volatile("" : "+r"(x)); }
template <typename T>
__device__ __forceinline__ void pin(T*& x) { asm volatile("" : "+l"(x)); }

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(GEMM_EPI_WARPS * 32) : "memory"); }

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
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// x / d for x * d < 2^32, d >= 1, with magic = ceil(2^32 / d) computed on the host (d = 1 -> magic 0 = "identity")
__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t magic) { return magic == 0u ? x : __umulhi(x, magic); }

struct TileCoord { int m_idx, n_idx, split; };
__device__ __forceinline__ TileCoord decode_tile(const GemmKernelParams& p, int tile) {
  const uint32_t split = fast_div(static_cast<uint32_t>(tile), p.magic_mn);
  const uint32_t mn = static_cast<uint32_t>(tile) - split * static_cast<uint32_t>(p.m_tiles * p.n_tiles);
  const uint32_t n_idx = fast_div(mn, p.magic_m);
  return {static_cast<int>(mn - n_idx * static_cast<uint32_t>(p.m_tiles)), static_cast<int>(n_idx), static_cast<int>(split)};
}

// One 16-column unit of the epilogue: v = acc * scale + bias (+ aux | * gelu'(aux)) -> activation -> staging box.  Kept small
// and called from a ROLLED loop on purpose: a tile's epilogue runs once, straight after a kernel switch, so every
// instruction line it touches is an instruction-cache miss; the unrolled 32-wide version spent ~3000 cycles per tile on
// instruction fetch alone (tools/gemm_trace.py).
__device__ __forceinline__ void epilogue_unit(const GemmKernelParams& p, const uint32_t (&r)[16], uint32_t s_bias, uint32_t s_scale,
                                              uint32_t xrow, uint32_t orow, int piece0, uint32_t swz64, uint32_t swz128,
                                              bool has_aux) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 s = lds_f4(s_scale + static_cast<uint32_t>(i) * 4u);
    const float4 b = lds_f4(s_bias + static_cast<uint32_t>(i) * 4u);
    v[i] = fmaf(__uint_as_float(r[i]), s.x, b.x);
    v[i + 1] = fmaf(__uint_as_float(r[i + 1]), s.y, b.y);
    v[i + 2] = fmaf(__uint_as_float(r[i + 2]), s.z, b.z);
    v[i + 3] = fmaf(__uint_as_float(r[i + 3]), s.w, b.w);
  }
  // bf16 boxes: 32 columns = four 16-byte pieces per 64-byte row (64B-swizzled); this unit is pieces piece0, piece0 + 1
  if (p.has_preact) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
      sts_u4(xrow + ((static_cast<uint32_t>(piece0 + j) ^ swz64) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]),
             pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
  }
  if (has_aux) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint4 a = lds_u4(xrow + ((static_cast<uint32_t>(piece0 + j) ^ swz64) << 4));
      const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
      const float av[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
      if (p.aux_mode == VB_AUX_ADD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * j + i] += av[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * j + i] *= gelu_fast_grad(av[i]);
      }
    }
  }
  if (p.act == VB_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = gelu_fast(v[i]);
  } else if (p.act == VB_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (p.act == VB_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = tanh_fast(v[i]);
  }
  if (p.d_is_f32) {
    // fp32 boxes: 32 columns = eight 16-byte pieces per 128-byte row (128B-swizzled); this unit is four of them
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts_u4(orow + ((static_cast<uint32_t>(2 * piece0 + j) ^ swz128) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
             __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
  } else {
#pragma unroll
    for (int j = 0; j < 2; ++j)
      sts_u4(orow + ((static_cast<uint32_t>(piece0 + j) ^ swz64) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]),
             pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
  }
}

__device__ __forceinline__ void epilogue_chunk(const GemmKernelParams& p, const uint32_t (&r)[32], int tcol, uint32_t s_bias,
                                               uint32_t s_scale, uint32_t xrow, uint32_t orow, uint32_t swz64, uint32_t swz128,
                                               bool has_aux) {
  uint32_t lo[16], hi[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { lo[i] = r[i]; hi[i] = r[16 + i]; }
  epilogue_unit(p, lo, s_bias + static_cast<uint32_t>(tcol) * 4u, s_scale + static_cast<uint32_t>(tcol) * 4u, xrow, orow, 0, swz64,
                swz128, has_aux);
  epilogue_unit(p, hi, s_bias + static_cast<uint32_t>(tcol + 16) * 4u, s_scale + static_cast<uint32_t>(tcol + 16) * 4u, xrow, orow, 2,
                swz64, swz128, has_aux);
}

// Epilogue data flow: the accumulator is read from TMEM one 32-column chunk per thread-row (both chunks a warp owns in a
// panel are fetched with one wait), combined with bias / scale (staged in smem), an optional aux panel (prefetched by
// TMA) and the activation, written into swizzled staging boxes (one box per chunk) and shipped with TMA stores (or TMA
// reduce-adds for split-K / accumulating fp32 outputs), so global traffic is fully coalesced and asynchronous.
template <int BN, bool A_MN, bool B_MN, int CG, int NP>
__global__ void __launch_bounds__(GEMM_THREADS, gemm_occupancy(BN))
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_x,
                 const GemmKernelParams p_const) {
  GemmKernelParams p = p_const;
  pin(p.scale); pin(p.bias); pin(p.m); pin(p.n); pin(p.k);
  pin(p.d_is_f32); pin(p.reduce_add); pin(p.act); pin(p.aux_mode); pin(p.has_preact);
  pin(p.splits); pin(p.kb_per_split); pin(p.m_tiles); pin(p.n_tiles); pin(p.stages); pin(p.out_bytes); pin(p.x_bytes);
  pin(p.magic_m); pin(p.magic_mn);
  asm volatile("" : "+l"(p.b_policy)); asm volatile("" : "+l"(p.d_policy));
  constexpr int BNL = BN / CG;                           // B rows (columns of the output) this CTA loads
  constexpr int CS = CG * NP;                            // CTAs per cluster: NP pairs, side by side along N, sharing A
  static_assert(CG == 2 || NP == 1, "multicast clusters are built from CTA pairs");
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BNL * GEMM_BK * 2;
  constexpr int NCHUNK = BN / GEMM_CHUNK;
  constexpr int GEMM_PANEL_CHUNKS = gemm_panel_chunks(BN);
  constexpr int PANELS = (NCHUNK + GEMM_PANEL_CHUNKS - 1) / GEMM_PANEL_CHUNKS;
  constexpr uint32_t TMEM_COLS = GEMM_ACC_STAGES * BN <= 128 ? 128 : (GEMM_ACC_STAGES * BN <= 256 ? 256 : 512);
  constexpr uint32_t IDESC = umma_idesc_bf16(GEMM_BM * CG, BN, A_MN, B_MN);
  static_assert(BN % GEMM_CHUNK == 0 && BN <= 256, "tile width");
  static_assert(!B_MN || BNL % 64 == 0, "an MN-major B tile is loaded in 64-wide pieces");
  static_assert(B_BYTES % 1024 == 0, "operand stages must keep the 1024-byte swizzle-atom alignment");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + STAGES * A_BYTES;
  uint8_t* stage_out = smem_b + STAGES * B_BYTES;
  uint8_t* stage_x = stage_out + p.out_bytes;
  float* s_bias = reinterpret_cast<float*>(stage_x + p.x_bytes);
  float* s_scale = s_bias + BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_scale + BN);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + GEMM_MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * GEMM_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + GEMM_ACC_STAGES;
  uint64_t* aux_full_bar = tmem_empty_bar + GEMM_ACC_STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(aux_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
  const uint32_t prank = rank & static_cast<uint32_t>(CG - 1);   // rank inside the pair: 0 = leader (issues the MMAs)
  const uint32_t pair = CG == 2 ? rank >> 1 : 0u;                // which pair of the cluster
  const uint32_t leader_rank = rank - prank;
  if (threadIdx.x == 0) trace_stamp(p, 0);
#ifdef VB_GEMM_TRACE
  if (threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); s_trace[22] = static_cast<long long>(g); }
#endif

  if (warp == 0) {
    // one barrier per lane, branch-free: [0,8) full (1: the leader's producer arrives, expecting the bytes of BOTH CTAs),
    // [8,16) empty (1: one multicast commit), 16/17 tmem_full (1), 18/19 tmem_empty (one arrive per epilogue warp of every
    // CTA of the pair), 20 aux (1)
    constexpr int NBARS = 2 * GEMM_MAX_STAGES + 2 * GEMM_ACC_STAGES + 1;
    const bool is_tmem_empty = lane >= 2 * GEMM_MAX_STAGES + GEMM_ACC_STAGES && lane < 2 * GEMM_MAX_STAGES + 2 * GEMM_ACC_STAGES;
    const bool is_empty = lane >= GEMM_MAX_STAGES && lane < 2 * GEMM_MAX_STAGES;   // one multicast commit per pair of the cluster
    if (lane < NBARS) mbar_init(&bars[lane], is_tmem_empty ? GEMM_EPI_WARPS * CG : (is_empty ? NP : 1));
    __syncwarp();
    if (lane == 0) {
      fence_mbar_init();
      trace_stamp(p, 19);
    }
  } else if (warp >= 2 && warp < 6 && lane == 0) {
    // descriptor prefetch, one per warp so that no warp serialises over several uniform-register operands
    if (warp == 2) tma_prefetch_desc(&tma_a);
    else if (warp == 3) tma_prefetch_desc(&tma_b);
    else if (warp == 4) tma_prefetch_desc(&tma_d);
    else if (p.x_bytes) tma_prefetch_desc(&tma_x);
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_pair(tmem_base_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_base_slot, TMEM_COLS);
      tmem_relinquish();
    }
    if (lane == 0) trace_stamp(p, 20);
  }
  tc_fence_before();
  __syncwarp();   // barrier.cluster is .aligned: warps must be converged
  if constexpr (CG == 2) {
    __syncthreads();            // CTA-scope ordering of the TMEM base slot / barrier inits
    if (threadIdx.x == 0) trace_stamp(p, 21);
    cluster_sync_relaxed();     // the peer's barriers exist (fence.mbarrier_init above is the cluster-scope release)
  } else {
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // PDL: the next kernel of the stream may begin its own prologue now (all our TMEM is allocated); we may not touch
  // global memory written by the previous kernel before it has completed.
  if (threadIdx.x == 0) trace_stamp(p, 1);
  griddep_launch();
  griddep_wait();
  if (threadIdx.x == 0) trace_stamp(p, 2);

  const int total_kb = (p.k + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = p.m_tiles * p.n_tiles * p.splits;
  const int first_tile = blockIdx.x / CS;
  const int tile_step = gridDim.x / CS;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one per CTA)
    if (lane == 0) {
      trace_stamp(p, 17);
      int stage = 0, fills = 0;
      uint32_t phase = 0;
      const uint32_t full_leader = CG == 2 ? mapa_u32(&full_bar[0], leader_rank) : 0u;
      const uint16_t a_mask = static_cast<uint16_t>((1u << prank) | (1u << (prank + 2)));   // me and my twin in the other pair
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const TileCoord tc = decode_tile(p, tile);
        const int m0 = tc.m_idx * (GEMM_BM * CG) + static_cast<int>(prank) * GEMM_BM;
        const int n0 = (tc.n_idx * NP + static_cast<int>(pair)) * BN + static_cast<int>(prank) * BNL;
        const int kb0 = tc.split * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (fills >= STAGES) mbar_wait(&empty_bar[stage], phase ^ 1u);   // the first pass over the ring needs no wait
          ++fills;
          if (tile == first_tile && kb == kb0) trace_stamp(p, 18);
#ifdef VB_GEMM_TRACE
          if (p.debug_mode == 2) {
            if (prank == 0) mbar_arrive(&full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
#endif
          uint8_t* sa = smem_a + stage * A_BYTES;
          uint8_t* sb = smem_b + stage * B_BYTES;
          const int k0 = kb * GEMM_BK;
          if constexpr (CG == 2) {
            // both CTAs' bytes are counted on the leader's barrier; a peer load that lands before the leader's
            // expect_tx only drives the transaction count negative for a moment (same phase: the peer cannot run ahead
            // of the leader's MMA, which frees the stage for both)
            const uint32_t bar = full_leader + static_cast<uint32_t>(stage) * 8u;
            if (prank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_BYTES + B_BYTES));
            if constexpr (NP == 2) {
              // the two pairs of the cluster work on the same 256 rows: each CTA fetches HALF of its A tile (64 rows) and
              // multicasts it to itself and to its twin in the other pair, which halves the A bytes read from L2.  Safe:
              // empty_bar counts the commits of BOTH pairs, so the twin's slot is free too.
              if constexpr (A_MN) tma_load_2d_pair_mc(sa + pair * (GEMM_BK * 128), &tma_a, bar, m0 + static_cast<int>(pair) * 64, k0, a_mask);
              else                tma_load_2d_pair_mc(sa + pair * (64 * 128), &tma_a, bar, k0, m0 + static_cast<int>(pair) * 64, a_mask);
            } else if constexpr (A_MN) {
#pragma unroll
              for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d_pair(sa + j * (GEMM_BK * 128), &tma_a, bar, m0 + j * 64, k0);
            } else {
              tma_load_2d_pair(sa, &tma_a, bar, k0, m0);
            }
            if constexpr (B_MN) {
#pragma unroll
              for (int j = 0; j < BNL / 64; ++j) {
                if (p.b_policy) tma_load_2d_pair_hint(sb + j * (GEMM_BK * 128), &tma_b, bar, n0 + j * 64, k0, p.b_policy);
                else            tma_load_2d_pair(sb + j * (GEMM_BK * 128), &tma_b, bar, n0 + j * 64, k0);
              }
            } else {
              if (p.b_policy) tma_load_2d_pair_hint(sb, &tma_b, bar, k0, n0, p.b_policy);
              else            tma_load_2d_pair(sb, &tma_b, bar, k0, n0);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
            if constexpr (A_MN) {
#pragma unroll
              for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d(sa + j * (GEMM_BK * 128), &tma_a, &full_bar[stage], m0 + j * 64, k0);
            } else {
              tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);
            }
            if constexpr (B_MN) {
#pragma unroll
              for (int j = 0; j < BNL / 64; ++j) {
                if (p.b_policy) tma_load_2d_hint(sb + j * (GEMM_BK * 128), &tma_b, &full_bar[stage], n0 + j * 64, k0, p.b_policy);
                else            tma_load_2d(sb + j * (GEMM_BK * 128), &tma_b, &full_bar[stage], n0 + j * 64, k0);
              }
            } else {
              if (p.b_policy) tma_load_2d_hint(sb, &tma_b, &full_bar[stage], k0, n0, p.b_policy);
              else            tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);
            }
          }
          if (tile == first_tile && kb == kb0) trace_stamp(p, 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      trace_stamp(p, 4);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (the leader CTA of every pair)
    if (prank == 0) {
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
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (tile == first_tile && kb == kb0 && lane == 0) trace_stamp(p, 5);
#ifdef VB_GEMM_TRACE
          if (p.debug_mode == 1) {
            if constexpr (CG == 2) umma_commit_pair_warp(&empty_bar[stage], all_mask); else umma_commit_warp(&empty_bar[stage]);
          } else
#endif
          {
            // descriptors differ from the stage base only in the 14-bit start-address field: +32 B per K step inside a
            // K-major swizzle row, +2048 B (16 k-rows) per K step of an MN-major tile
            const uint64_t da0 = A_MN ? umma_smem_desc(smem_u32(smem_a + stage * A_BYTES), GEMM_BK * 128, 1024)
                                      : umma_smem_desc(smem_u32(smem_a + stage * A_BYTES), 16, 1024);
            const uint64_t db0 = B_MN ? umma_smem_desc(smem_u32(smem_b + stage * B_BYTES), GEMM_BK * 128, 1024)
                                      : umma_smem_desc(smem_u32(smem_b + stage * B_BYTES), 16, 1024);
#pragma unroll
            for (int kk = 0; kk < GEMM_BK / 16; ++kk) {
              const uint64_t da = da0 + static_cast<uint64_t>(kk * (A_MN ? 128 : 2));
              const uint64_t db = db0 + static_cast<uint64_t>(kk * (B_MN ? 128 : 2));
              const uint32_t accum = (kb > kb0 || kk > 0) ? 1u : 0u;
              if constexpr (CG == 2) umma_bf16_pair_warp(tmem_d, da, db, IDESC, accum);
              else                   umma_bf16_warp(tmem_d, da, db, IDESC, accum);
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
      if (lane == 0) trace_stamp(p, 6);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9 of every CTA)
    const int quarter = warp & 3;           // TMEM sub-partition this warp may read: lanes [32q, 32q+32)
    const int half = (warp - 2) >> 2;       // even / odd chunks of a panel
    const int epi_tid = threadIdx.x - 64;
    const int ewarp = warp - 2;
    const bool leader = epi_tid == 0;
    const int row = quarter * 32 + lane;    // row inside this CTA's 128-row block
    const uint32_t swz128 = static_cast<uint32_t>(row & 7);          // 128-byte rows (fp32 boxes)
    const uint32_t swz64 = static_cast<uint32_t>((row >> 1) & 3);    // 64-byte rows (bf16 boxes)
    const bool has_aux = p.aux_mode != VB_AUX_NONE;
    const uint32_t tmem_empty_leader = CG == 2 ? mapa_u32(&tmem_empty_bar[0], leader_rank) : 0u;
    const uint32_t sa_bias = smem_u32(s_bias), sa_scale = smem_u32(s_scale);
    const uint32_t sa_x = smem_u32(stage_x) + static_cast<uint32_t>(row) * 64u;
    const uint32_t sa_o16 = smem_u32(stage_out) + static_cast<uint32_t>(row) * 64u;
    const uint32_t sa_o32 = smem_u32(stage_out) + static_cast<uint32_t>(row) * 128u;
    int acc = 0;
    uint32_t acc_phase = 0, aux_phase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile(p, tile);
      const int m0 = tc.m_idx * (GEMM_BM * CG) + static_cast<int>(prank) * GEMM_BM;
      const int n0 = (tc.n_idx * NP + static_cast<int>(pair)) * BN;
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int panel = 0; panel < PANELS; ++panel) {
        const int pn0 = n0 + panel * (GEMM_PANEL_CHUNKS * GEMM_CHUNK);
        const int nch = min(GEMM_PANEL_CHUNKS, NCHUNK - panel * GEMM_PANEL_CHUNKS);
        if (lane == 0 && ewarp < GEMM_PANEL_CHUNKS) tma_store_wait_read();   // my staging box of the previous panel has been read
        if (leader) {
          if (has_aux) {
            mbar_arrive_expect_tx(aux_full_bar, static_cast<uint32_t>(nch) * GEMM_BOX_BF16);
            for (int b = 0; b < nch; ++b)
              tma_load_2d(stage_x + b * GEMM_BOX_BF16, &tma_x, aux_full_bar, pn0 + b * GEMM_CHUNK, m0);
          }
        }
        if (panel == 0) {
          for (int i = epi_tid; i < BN; i += GEMM_EPI_WARPS * 32) {
            const bool ok = n0 + i < p.n;
            s_bias[i] = (p.bias != nullptr && ok) ? __ldg(p.bias + n0 + i) : 0.0f;
            s_scale[i] = (p.scale != nullptr && ok) ? __ldg(p.scale + n0 + i) : 1.0f;
          }
        }
        epi_barrier();                      // staging free, bias / scale visible
        if (panel == 0) {
          mbar_wait(&tmem_full_bar[acc], acc_phase);
          tc_fence_after();
          if (leader && tile == first_tile) trace_stamp(p, 7);
          if (leader) trace_stamp(p, 8);     // last tile's accumulator ready
        }
        // this warp owns chunks `half` and `half + 2` of the panel: fetch both from TMEM, wait once
        const int c_a = half, c_b = half + 2;
        const bool do_a = c_a < nch, do_b = c_b < nch;
        uint32_t r_a[32], r_b[GEMM_PANEL_CHUNKS > 2 ? 32 : 1];
        if (do_a) tmem_ld_32x32(taddr + static_cast<uint32_t>((panel * GEMM_PANEL_CHUNKS + c_a) * GEMM_CHUNK), r_a);
        if constexpr (GEMM_PANEL_CHUNKS > 2) {
          if (do_b) tmem_ld_32x32(taddr + static_cast<uint32_t>((panel * GEMM_PANEL_CHUNKS + c_b) * GEMM_CHUNK), r_b);
        }
        if (has_aux) {
          mbar_wait(aux_full_bar, aux_phase);
          aux_phase ^= 1u;
        }
        tmem_ld_wait();
        if (leader) trace_stamp(p, 12);
        if (panel == PANELS - 1) {
          // accumulator stage drained into registers: hand it back to the MMA warp (of the leader) right away
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster(tmem_empty_leader + static_cast<uint32_t>(acc) * 8u);
            else                   mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        if (do_a)
          epilogue_chunk(p, r_a, (panel * GEMM_PANEL_CHUNKS + c_a) * GEMM_CHUNK, sa_bias, sa_scale, sa_x + c_a * GEMM_BOX_BF16,
                         p.d_is_f32 ? sa_o32 + c_a * GEMM_BOX_F32 : sa_o16 + c_a * GEMM_BOX_BF16, swz64, swz128, has_aux);
        if constexpr (GEMM_PANEL_CHUNKS > 2) {
          if (do_b)
            epilogue_chunk(p, r_b, (panel * GEMM_PANEL_CHUNKS + c_b) * GEMM_CHUNK, sa_bias, sa_scale, sa_x + c_b * GEMM_BOX_BF16,
                           p.d_is_f32 ? sa_o32 + c_b * GEMM_BOX_F32 : sa_o16 + c_b * GEMM_BOX_BF16, swz64, swz128, has_aux);
        }
        if (leader) trace_stamp(p, 13);
        fence_proxy_async_smem();           // make the staged panel visible to the TMA engine
        if (leader) trace_stamp(p, 14);
        epi_barrier();
        // one TMA store per staging box, issued by the first lane of epilogue warp `box` (a UTMASTG issue costs ~190 cycles
        // in the issuing thread: eight of them back to back would serialise)
        if (lane == 0 && ewarp < nch) {
          if (leader) trace_stamp(p, 15);
          const int c0 = pn0 + ewarp * GEMM_CHUNK;
          if (c0 < p.n) {
            if (p.d_is_f32) {
              if (p.d_policy) {
                if (p.reduce_add) tma_reduce_add_2d_hint(&tma_d, stage_out + ewarp * GEMM_BOX_F32, c0, m0, p.d_policy);
                else              tma_store_2d_hint(&tma_d, stage_out + ewarp * GEMM_BOX_F32, c0, m0, p.d_policy);
              } else {
                if (p.reduce_add) tma_reduce_add_2d(&tma_d, stage_out + ewarp * GEMM_BOX_F32, c0, m0);
                else              tma_store_2d(&tma_d, stage_out + ewarp * GEMM_BOX_F32, c0, m0);
              }
            } else {
              tma_store_2d(&tma_d, stage_out + ewarp * GEMM_BOX_BF16, c0, m0);
              if (p.has_preact) tma_store_2d(&tma_x, stage_x + ewarp * GEMM_BOX_BF16, c0, m0);
            }
          }
          tma_store_commit();
          if (leader) trace_stamp(p, 16);
        }
      }
      if (++acc == GEMM_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
    if (leader) trace_stamp(p, 9);
    if (lane == 0 && ewarp < GEMM_PANEL_CHUNKS) tma_store_wait_read();   // smem may be released; the writes themselves complete before the grid does
    if (leader) trace_stamp(p, 10);
  }

  tc_fence_before();
  __syncwarp();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_relaxed();   // the other CTAs' smem / barriers stay alive until all are done
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
#ifdef VB_GEMM_TRACE
  if (threadIdx.x == 0) trace_stamp(p, 11);
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

// Co-resident clusters of `cs` CTAs (one CTA per SM).  Clusters cannot straddle a GPC and the B200's GPCs do not all hold a
// multiple of four SMs, so fewer than 148 / cs clusters of four fit; asked of the driver once per cluster size.
static int max_clusters(int cs, int occupancy);

template <int BN, bool A_MN, bool B_MN, int CG, int NP>
static int launch_gemm(const vb_gemm_args& a, int splits, cudaStream_t stream) {
  constexpr int BNL = BN / CG;
  constexpr int CS = CG * NP;
  constexpr int STAGE_BYTES = (GEMM_BM + BNL) * GEMM_BK * 2;
  constexpr int NCHUNK = BN / GEMM_CHUNK;
  constexpr int GEMM_PANEL_CHUNKS = gemm_panel_chunks(BN);
  constexpr int GEMM_SMEM_LIMIT = gemm_smem_limit(BN);
  constexpr int PANEL = NCHUNK < GEMM_PANEL_CHUNKS ? NCHUNK : GEMM_PANEL_CHUNKS;
  CUtensorMap map_a, map_b, map_d, map_x;
  int rc;
  // K-major operand: global [rows, K] -> box {64 (k), rows_per_cta}; MN-major: global [K, rows] -> box {64 (mn), 64 (k)}
  if (A_MN) rc = make_tensor_map_2d(&map_a, a.a, /*inner*/ a.m, /*outer*/ a.k, a.lda, 64, GEMM_BK);
  else      rc = make_tensor_map_2d(&map_a, a.a, a.k, a.m, a.lda, GEMM_BK, GEMM_BM / NP);   // NP = 2: multicast halves
  if (rc != VB_OK) return rc;
  if (B_MN) rc = make_tensor_map_2d(&map_b, a.b, a.n, a.k, a.ldb, 64, GEMM_BK);
  else      rc = make_tensor_map_2d(&map_b, a.b, a.k, a.n, a.ldb, GEMM_BK, BNL);
  if (rc != VB_OK) return rc;
  // epilogue boxes: 128 rows x 32 columns (fp32: 128-byte rows, 128B swizzle; bf16: 64-byte rows, 64B swizzle)
  if (a.d_is_f32) rc = make_tensor_map_2d_f32(&map_d, a.d, a.n, a.m, a.ldd, GEMM_CHUNK, GEMM_BM);
  else            rc = make_tensor_map_2d_sw64(&map_d, a.d, a.n, a.m, a.ldd, GEMM_CHUNK, GEMM_BM);
  if (rc != VB_OK) return rc;
  const void* xptr = a.d_preact != nullptr ? a.d_preact : a.aux;
  const int64_t ldx = a.d_preact != nullptr ? a.ld_preact : a.ld_aux;
  if (xptr != nullptr) {
    rc = make_tensor_map_2d_sw64(&map_x, xptr, a.n, a.m, ldx, GEMM_CHUNK, GEMM_BM);
    if (rc != VB_OK) return rc;
  } else {
    map_x = map_d;
  }

  GemmKernelParams p;
  p.scale = a.scale; p.bias = a.bias;
  p.m = a.m; p.n = a.n; p.k = a.k;
  p.d_is_f32 = a.d_is_f32; p.act = a.act; p.aux_mode = a.aux_mode;
  p.has_preact = a.d_preact != nullptr;
  p.m_tiles = (a.m + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  p.n_tiles = ((a.n + BN - 1) / BN + NP - 1) / NP;   // N steps of a whole cluster (NP tiles side by side)
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
  p.out_bytes = PANEL * (a.d_is_f32 ? GEMM_BOX_F32 : GEMM_BOX_BF16);
  p.x_bytes = xptr != nullptr ? PANEL * GEMM_BOX_BF16 : 0;
  const int fixed = 1024 /*alignment slack*/ + p.out_bytes + p.x_bytes + 2 * BN * 4 + 256 /*barriers*/;
  int stages = (GEMM_SMEM_LIMIT - fixed) / STAGE_BYTES;
  if (stages > GEMM_MAX_STAGES) stages = GEMM_MAX_STAGES;
  if (stages < 2) {
    vb_set_last_error("vb_gemm_bf16", "tile configuration does not fit shared memory");
    return VB_ERR_UNSUPPORTED;
  }
  static const int debug_mode = getenv("VB_GEMM_DEBUG") ? atoi(getenv("VB_GEMM_DEBUG")) : 0;
  static const int debug_stages = getenv("VB_GEMM_STAGES") ? atoi(getenv("VB_GEMM_STAGES")) : 0;
  p.debug_mode = debug_mode;
  p.trace = g_trace;
  static const bool hints = !(getenv("VB_GEMM_NO_L2_HINTS") && atoi(getenv("VB_GEMM_NO_L2_HINTS")) != 0);
  p.b_policy = (hints && a.b_streamed) ? L2_EVICT_FIRST : 0ull;     // 0 = the plain (un-hinted) instruction
  p.d_policy = (hints && a.d_streamed && a.d_is_f32) ? L2_EVICT_FIRST : 0ull;
  if (debug_stages >= 2 && debug_stages < stages) stages = debug_stages;
  p.stages = stages;
  const int smem_bytes = fixed + stages * STAGE_BYTES;

  static bool attr_set = false;   // one per instantiation
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, CG, NP>;
  if (!attr_set) {
    VB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
    attr_set = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.splits;
  int groups = max_clusters(CS, gemm_occupancy(BN));
  if (a.max_ctas > 0 && a.max_ctas / CS < groups) groups = a.max_ctas / CS > 0 ? a.max_ctas / CS : 1;
  if (tiles < groups) groups = tiles;
  VB_CUDA_CHECK(launch_ex(kern, dim3(groups * CS), dim3(GEMM_THREADS), smem_bytes, stream, CS, /*pdl=*/true, map_a, map_b,
                          map_d, map_x, p));
  return VB_OK;
}

static int max_clusters(int cs, int occupancy) {
  static int cached[5][3] = {};
  if (cs <= 1) return num_sms() * occupancy;
  if (cached[cs][occupancy] == 0) {
    int n = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(num_sms() / cs * cs));
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = occupancy == 2 ? gemm_smem_limit(128) : gemm_smem_limit(256);
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
        auto kern = gemm_bf16_kernel<128, false, false, 2, 1>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_limit(128));
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      } else {
        auto kern = gemm_bf16_kernel<128, false, false, 2, 2>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_limit(128));
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      }
    } else {
      if (cs == 2) {
        auto kern = gemm_bf16_kernel<256, false, false, 2, 1>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_limit(256));
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      } else {
        auto kern = gemm_bf16_kernel<256, false, false, 2, 2>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_limit(256));
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      }
    }
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms() / cs * 7 / 8 * occupancy; }
    cached[cs][occupancy] = n;
  }
  return cached[cs][occupancy];
}

// Tile-shape heuristic.  Cost of a tile in SM cycles = k-blocks x max(tensor-pipe floor, what the SM can ingest, its share
// of what L2 can deliver) + epilogue; the candidate with the cheapest (waves x tile cost) wins.  Constants measured with
// tools/gemm_trace.py / gemm_ksweep.py on B200: L2 delivers ~6300 B/clk to all SMs together, one SM takes ~58 B/clk, a
// UMMA 128 x N x 16 (per SM) costs max(~110, N/2 + 20) cycles.
struct TileChoice { int bn, cg, np, splits; };

static double tile_cost(int bn, int cg, int np, int kb, double ctas, bool f32_out, bool split) {
  // `ctas` CTAs run at once; when that is more than one per SM (two small CTAs co-resident) they share the SM's ingest
  // bandwidth and tensor pipe
  const double sms = ctas < num_sms() ? ctas : num_sms();
  const double share = ctas / sms;
  const double a_bytes = GEMM_BM * 128.0, b_bytes = (bn / cg) * 128.0;
  const double ingest = (a_bytes + b_bytes) / 58.0 * share;
  const double l2 = (a_bytes / np + b_bytes) * ctas / 6300.0;
  const double mma_one = bn / (2.0 * (cg == 2 ? 1.0 : 1.0)) + 20.0;
  const double mma = 4.0 * (mma_one < 110.0 ? 110.0 : mma_one) * share;
  double kb_cost = ingest > mma ? ingest : mma;
  if (l2 > kb_cost) kb_cost = l2;
  const double epi = 900.0 + (f32_out ? 6.0 : 4.0) * bn + (split ? 2.0 * bn : 0.0);
  return kb * kb_cost + epi;
}

static bool tile_legal(const vb_gemm_args& a, int bn, int cg) {
  static const int max_bn = getenv("VB_GEMM_MAX_BN") ? atoi(getenv("VB_GEMM_MAX_BN")) : 256;   // experiments: compact tiles only
  if (bn > max_bn) return false;
  if (cg == 1) return bn == 64 || bn == 128;
  if (a.b_mn_major) return bn == 128 || bn == 256;              // 64-wide MN pieces per CTA
  if (a.a_mn_major) return bn == 128;                           // (MN, K): API completeness only
  return bn == 64 || bn == 96 || bn == 128 || bn == 192 || bn == 256;
}

static TileChoice pick_config(const vb_gemm_args& a) {
  static const int force_np = getenv("VB_GEMM_NP") ? atoi(getenv("VB_GEMM_NP")) : 0;
  static const int force_cg = getenv("VB_GEMM_CG") ? atoi(getenv("VB_GEMM_CG")) : 0;
  const int total_kb = (a.k + GEMM_BK - 1) / GEMM_BK;
  const bool can_split = a.d_is_f32 && a.accumulate;
  double best_cost = 1e30;
  TileChoice best = {128, a.m > GEMM_BM ? 2 : 1, 1, 1};
  const int bns[5] = {128, 96, 192, 256, 64};
  for (int pass = 0; pass < 2 && best_cost > 1e29; ++pass) {     // pass 1: the block_n hint was not legal, ignore it
    for (int cg = 2; cg >= 1; --cg) {
      if (cg == 2 && a.m <= GEMM_BM) continue;                   // a pair would leave its second CTA without rows
      if (force_cg != 0 && a.m > GEMM_BM && cg != force_cg) continue;
      const int m_tiles = (a.m + GEMM_BM * cg - 1) / (GEMM_BM * cg);
      for (int np = 1; np <= cg; ++np) {
        if (force_np != 0 && cg == 2 && np != force_np) continue;
        for (int bi = 0; bi < 5; ++bi) {
          const int bn = bns[bi];
          if (!tile_legal(a, bn, cg)) continue;
          int clusters = max_clusters(cg * np, gemm_occupancy(bn));
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
            const double cost = waves * tile_cost(bn, cg, np, kb, ctas, a.d_is_f32 != 0, s > 1);
            if (cost < best_cost * 0.97) { best_cost = cost; best = {bn, cg, np, s}; }
          }
        }
      }
    }
  }
  return best;
}

template <int BN, int CG, int NP>
static int dispatch_major(const vb_gemm_args& a, int splits, cudaStream_t s) {
  if (a.a_mn_major && a.b_mn_major) {
    if constexpr (BN == 128 || BN == 256 || CG == 1) return launch_gemm<BN, true, true, CG, NP>(a, splits, s);
  } else if (a.a_mn_major) {
    if constexpr (BN == 128 || CG == 1) return launch_gemm<BN, true, false, CG, NP>(a, splits, s);
  } else if (a.b_mn_major) {
    if constexpr (BN == 128 || BN == 256 || CG == 1) return launch_gemm<BN, false, true, CG, NP>(a, splits, s);
  } else {
    return launch_gemm<BN, false, false, CG, NP>(a, splits, s);
  }
  vb_set_last_error("vb_gemm_bf16", "internal: tile shape not instantiated for this operand layout");
  return VB_ERR_UNSUPPORTED;
}

template <int NP>
static int dispatch_bn(const vb_gemm_args& a, const TileChoice& c, cudaStream_t s) {
  switch (c.bn) {
    case 64: return dispatch_major<64, 2, NP>(a, c.splits, s);
    case 96: return dispatch_major<96, 2, NP>(a, c.splits, s);
    case 192: return dispatch_major<192, 2, NP>(a, c.splits, s);
    case 256: return dispatch_major<256, 2, NP>(a, c.splits, s);
    default: return dispatch_major<128, 2, NP>(a, c.splits, s);
  }
}

}  // namespace vb

// diagnostic: resident blocks per SM / co-resident 2-CTA clusters the runtime reports for the narrow-tile pair kernel at a given
// dynamic shared-memory size (tools/gemm_occupancy.py)
extern "C" int vb_gemm_debug_occupancy(int smem_bytes, int* blocks_per_sm, int* clusters) {
  auto kern = vb::gemm_bf16_kernel<128, false, false, 2, 1>;
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
  VB_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "lda/ldb must be multiples of 8 elements (16-byte TMA strides)");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(a.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.b) & 15) == 0,
             "a/b must be 16-byte aligned");
  VB_REQUIRE(a.ldd % (a.d_is_f32 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(a.d) & 15) == 0, "d must be 16-byte aligned rows");
  VB_REQUIRE(a.d_preact == nullptr || (a.ld_preact % 8 == 0 && (reinterpret_cast<uintptr_t>(a.d_preact) & 15) == 0), "d_preact alignment");
  VB_REQUIRE(a.aux_mode == VB_AUX_NONE || (a.aux != nullptr && a.ld_aux % 8 == 0 && (reinterpret_cast<uintptr_t>(a.aux) & 15) == 0), "aux missing or misaligned");
  VB_REQUIRE(a.scale == nullptr || (reinterpret_cast<uintptr_t>(a.scale) & 15) == 0, "scale alignment");
  VB_REQUIRE(a.bias == nullptr || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "bias alignment");
  VB_REQUIRE(a.block_n == 0 || a.block_n == 64 || a.block_n == 96 || a.block_n == 128 || a.block_n == 192 || a.block_n == 256,
             "block_n must be 0, 64, 96, 128, 192 or 256");
  VB_REQUIRE(a.splits >= 0 && (a.splits <= 1 || (a.d_is_f32 && a.accumulate)), "split-K needs an fp32 accumulating output");
  VB_REQUIRE(!(a.accumulate && !a.d_is_f32), "accumulate needs an fp32 output");
  VB_REQUIRE(!(a.d_is_f32 && a.d_preact != nullptr), "d_preact only with a bf16 output");
  VB_REQUIRE(!(a.d_preact != nullptr && a.aux_mode != VB_AUX_NONE), "d_preact and aux share one staging panel: use one of them");
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
    fprintf(stderr, "vb_gemm %dx%dx%d a_mn=%d b_mn=%d f32=%d -> bn=%d cg=%d np=%d splits=%d (clusters %d)\n", a.m, a.n, a.k, a.a_mn_major,
            a.b_mn_major, a.d_is_f32, c.bn, c.cg, c.np, c.splits, max_clusters(c.cg * c.np, gemm_occupancy(c.bn)));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (c.cg == 1) {
    if (c.bn == 64) return dispatch_major<64, 1, 1>(a, c.splits, s);
    return dispatch_major<128, 1, 1>(a, c.splits, s);
  }
  if (c.np == 2) return dispatch_bn<2>(a, c, s);
  return dispatch_bn<1>(a, c, s);
}
