// tcgen05 / TMEM GEMM for sm_100a, fed by TMA.  One persistent, warp-specialised kernel:
//
//   warp 0      TMA producer      global -> 128B-swizzled smem ring (mbarrier full/empty)
//   warp 1      MMA issuer        one thread issues tcgen05.mma (UMMA 128 x BN x 16, bf16 -> fp32 in TMEM)
//   warps 2..5  epilogue          tcgen05.ld accumulator -> scale/bias/aux/activation -> global
//
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// Operands may be K-major or MN-major (descriptor + TMA box change only), so forward, dgrad and
// wgrad of nn.Linear (reference models/vilbert_facebook_arch.py:127-129 etc.) all run here without a
// transposed copy of anything.  See include/vilbert_b200.h for the ABI.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/vilbert_b200.h"
#include "tensormap.h"

namespace vb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;         // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_EPI_WARPS = 8;   // two per TMEM sub-partition, each takes half of a column panel
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_ACC_STAGES = 2;
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_BOX_BYTES = GEMM_BM * 128;  // one epilogue staging box: 128 rows x 128 bytes, 128B-swizzled
constexpr int GEMM_SMEM_LIMIT = 227 * 1024;

struct GemmKernelParams {
  const float* scale;
  const float* bias;
  int m, n, k;
  int d_is_f32, reduce_add, act, aux_mode, has_preact;
  int splits, kb_per_split;
  int m_tiles, n_tiles;
  int stages;       // depth of the operand ring
  int out_bytes;    // staging bytes for one column panel of the output
  int x_bytes;      // staging bytes for the aux-in / preact-out panel (0 = unused)
  int debug_mode;   // profiling only (VB_GEMM_DEBUG): 1 = no MMA issue, 2 = no TMA loads; results are garbage
};


__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(GEMM_EPI_WARPS * 32) : "memory"); }

// Epilogue data flow: the accumulator is read from TMEM one 32-column chunk per thread-row, combined with bias / scale
// (staged in smem), an optional aux panel (prefetched by TMA while the main loop of the tile is still running) and the
// activation, written into 128B-swizzled staging boxes and shipped with TMA stores (or TMA reduce-adds for split-K /
// accumulating fp32 outputs), so global traffic is fully coalesced and asynchronous.
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_x,
                 const GemmKernelParams p) {
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr int PN = BN < 128 ? BN : 128;                // epilogue column panel
  constexpr int PANELS = BN / PN;
  constexpr uint32_t TMEM_COLS = GEMM_ACC_STAGES * BN;   // 128, 256 or 512: all powers of two >= 32
  constexpr uint32_t IDESC = umma_idesc_bf16(GEMM_BM, BN, A_MN, B_MN);

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

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_d);
    if (p.x_bytes) tma_prefetch_desc(&tma_x);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < GEMM_ACC_STAGES; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], GEMM_EPI_WARPS);  // one arrive per epilogue warp
    }
    mbar_init(aux_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int total_kb = (p.k + GEMM_BK - 1) / GEMM_BK;
  const int mn_tiles = p.m_tiles * p.n_tiles;
  const int num_tiles = mn_tiles * p.splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mn = tile % mn_tiles;
        const int split = tile / mn_tiles;
        const int m0 = (mn % p.m_tiles) * GEMM_BM;
        const int n0 = (mn / p.m_tiles) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (p.debug_mode == 2 || p.debug_mode == 4) {
            mbar_arrive(&full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
          uint8_t* sa = smem_a + stage * A_BYTES;
          uint8_t* sb = smem_b + stage * B_BYTES;
          const int k0 = kb * GEMM_BK;
          if constexpr (A_MN) {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d(sa + j * (GEMM_BK * 128), &tma_a, &full_bar[stage], m0 + j * 64, k0);
          } else {
            tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (GEMM_BK * 128), &tma_b, &full_bar[stage], n0 + j * 64, k0);
          } else {
            tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int split = tile / mn_tiles;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(total_kb, kb0 + p.kb_per_split);
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (p.debug_mode == 1) {
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
        } else {
          // descriptors differ from the stage base only in the 14-bit start-address field: +32 B per K step inside a
          // K-major swizzle row, +2048 B (16 k-rows) per K step of an MN-major tile
          const uint64_t da0 = A_MN ? umma_smem_desc(smem_u32(smem_a + stage * A_BYTES), GEMM_BK * 128, 1024)
                                    : umma_smem_desc(smem_u32(smem_a + stage * A_BYTES), 16, 1024);
          const uint64_t db0 = B_MN ? umma_smem_desc(smem_u32(smem_b + stage * B_BYTES), GEMM_BK * 128, 1024)
                                    : umma_smem_desc(smem_u32(smem_b + stage * B_BYTES), 16, 1024);
#pragma unroll
          for (int kk = 0; kk < GEMM_BK / 16; ++kk)
            umma_bf16_warp(tmem_d + ((p.debug_mode >= 3 && (kk & 1)) ? BN : 0), da0 + static_cast<uint64_t>(kk * (A_MN ? 128 : 2)),
                           db0 + static_cast<uint64_t>(kk * (B_MN ? 128 : 2)), IDESC, (kb > kb0 || kk > 0) ? 1u : 0u);
          umma_commit_warp(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (p.debug_mode == 1) {
        if (lane == 0) mbar_arrive(&tmem_full_bar[acc]);
      } else {
        umma_commit_warp(&tmem_full_bar[acc]);
      }
      __syncwarp();
      if (++acc == GEMM_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int quarter = warp & 3;           // TMEM sub-partition this warp may read: lanes [32q, 32q+32)
    const int half = (warp - 2) >> 2;       // which half of a column panel this warp converts
    const int epi_tid = threadIdx.x - 64;
    const bool leader = epi_tid == 0;
    const int row = quarter * 32 + lane;    // row inside the tile
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const bool has_aux = p.aux_mode != VB_AUX_NONE;
    int acc = 0;
    uint32_t acc_phase = 0, aux_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mn = tile % mn_tiles;
      const int m0 = (mn % p.m_tiles) * GEMM_BM;
      const int n0 = (mn / p.m_tiles) * BN;
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int panel = 0; panel < PANELS; ++panel) {
        const int pn0 = n0 + panel * PN;
        if (leader) {
          tma_store_wait_read();            // staging boxes of the previous panel have been read by the TMA engine
          if (has_aux) {
            mbar_arrive_expect_tx(aux_full_bar, GEMM_BM * PN * 2);
#pragma unroll
            for (int b = 0; b < (PN + 63) / 64; ++b)
              tma_load_2d(stage_x + b * GEMM_BOX_BYTES, &tma_x, aux_full_bar, pn0 + b * 64, m0);
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
        }
        if (has_aux) {
          mbar_wait(aux_full_bar, aux_phase);
          aux_phase ^= 1u;
        }
#pragma unroll 1
        for (int cc = 0; cc < PN / 64; ++cc) {
          const int pcol = half * (PN / 2) + cc * 32;   // first of my 32 columns inside the panel
          const int tcol = panel * PN + pcol;           // ... inside the tile
          uint32_t r[32];
          tmem_ld_32x32(taddr + tcol, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 s = *reinterpret_cast<const float4*>(s_scale + tcol + i);
            const float4 b = *reinterpret_cast<const float4*>(s_bias + tcol + i);
            v[i] = fmaf(__uint_as_float(r[i]), s.x, b.x);
            v[i + 1] = fmaf(__uint_as_float(r[i + 1]), s.y, b.y);
            v[i + 2] = fmaf(__uint_as_float(r[i + 2]), s.z, b.z);
            v[i + 3] = fmaf(__uint_as_float(r[i + 3]), s.w, b.w);
          }
          // bf16 panels: 64 columns per box, my 32 columns are 16-byte pieces [piece0, piece0+4) of box `xbox`
          const int xbox = pcol >> 6;
          const uint32_t piece0 = static_cast<uint32_t>((pcol & 63) >> 3);
          uint8_t* xrow = stage_x + xbox * GEMM_BOX_BYTES + row * 128;
          if (p.has_preact) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * j], v[8 * j + 1]); o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(xrow + (((piece0 + j) ^ swz) << 4)) = o;
            }
          }
          if (has_aux) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 a = *reinterpret_cast<const uint4*>(xrow + (((piece0 + j) ^ swz) << 4));
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
            for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
          } else if (p.act == VB_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
          } else if (p.act == VB_ACT_TANH) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
          }
          if (p.d_is_f32) {
            // fp32 panels: 32 columns per box = exactly my chunk, eight 16-byte pieces
            uint8_t* orow = stage_out + (pcol >> 5) * GEMM_BOX_BYTES + row * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(orow + ((static_cast<uint32_t>(j) ^ swz) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint8_t* orow = stage_out + xbox * GEMM_BOX_BYTES + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * j], v[8 * j + 1]); o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(orow + (((piece0 + j) ^ swz) << 4)) = o;
            }
          }
        }
        if (panel == PANELS - 1) {
          // accumulator stage drained: hand it back to the MMA warp before the stores are even issued
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        fence_proxy_async_smem();           // make the staged panel visible to the TMA engine
        epi_barrier();
        if (leader && pn0 < p.n) {
          if (p.d_is_f32) {
#pragma unroll
            for (int b = 0; b < PN / 32; ++b) {
              if (pn0 + b * 32 < p.n) {
                if (p.reduce_add) tma_reduce_add_2d(&tma_d, stage_out + b * GEMM_BOX_BYTES, pn0 + b * 32, m0);
                else              tma_store_2d(&tma_d, stage_out + b * GEMM_BOX_BYTES, pn0 + b * 32, m0);
              }
            }
          } else {
#pragma unroll
            for (int b = 0; b < (PN + 63) / 64; ++b) {
              if (pn0 + b * 64 < p.n) {
                tma_store_2d(&tma_d, stage_out + b * GEMM_BOX_BYTES, pn0 + b * 64, m0);
                if (p.has_preact) tma_store_2d(&tma_x, stage_x + b * GEMM_BOX_BYTES, pn0 + b * 64, m0);
              }
            }
          }
          tma_store_commit();
        }
      }
      if (++acc == GEMM_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
    if (leader) tma_store_wait_read();   // smem may be released; the writes themselves complete before the grid does
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static int g_num_sms = 0;

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

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const vb_gemm_args& a, int splits, cudaStream_t stream) {
  constexpr int PN = BN < 128 ? BN : 128;
  constexpr int STAGE_BYTES = (GEMM_BM + BN) * GEMM_BK * 2;
  CUtensorMap map_a, map_b, map_d, map_x;
  int rc;
  // K-major operand: global [rows, K] -> box {64 (k), rows_per_tile}; MN-major: global [K, rows] -> box {64 (mn), 64 (k)}
  if (A_MN) rc = make_tensor_map_2d(&map_a, a.a, /*inner*/ a.m, /*outer*/ a.k, a.lda, 64, GEMM_BK);
  else      rc = make_tensor_map_2d(&map_a, a.a, a.k, a.m, a.lda, GEMM_BK, GEMM_BM);
  if (rc != VB_OK) return rc;
  if (B_MN) rc = make_tensor_map_2d(&map_b, a.b, a.n, a.k, a.ldb, 64, GEMM_BK);
  else      rc = make_tensor_map_2d(&map_b, a.b, a.k, a.n, a.ldb, GEMM_BK, BN);
  if (rc != VB_OK) return rc;
  // epilogue boxes: 128 rows x 128 bytes (64 bf16 or 32 fp32 columns)
  if (a.d_is_f32) rc = make_tensor_map_2d_f32(&map_d, a.d, a.n, a.m, a.ldd, 32, GEMM_BM);
  else            rc = make_tensor_map_2d(&map_d, a.d, a.n, a.m, a.ldd, PN < 64 ? PN : 64, GEMM_BM);
  if (rc != VB_OK) return rc;
  const void* xptr = a.d_preact != nullptr ? a.d_preact : a.aux;
  const int64_t ldx = a.d_preact != nullptr ? a.ld_preact : a.ld_aux;
  if (xptr != nullptr) {
    rc = make_tensor_map_2d(&map_x, xptr, a.n, a.m, ldx, PN < 64 ? PN : 64, GEMM_BM);
    if (rc != VB_OK) return rc;
  } else {
    map_x = map_d;
  }

  GemmKernelParams p;
  p.scale = a.scale; p.bias = a.bias;
  p.m = a.m; p.n = a.n; p.k = a.k;
  p.d_is_f32 = a.d_is_f32; p.act = a.act; p.aux_mode = a.aux_mode;
  p.has_preact = a.d_preact != nullptr;
  p.m_tiles = (a.m + GEMM_BM - 1) / GEMM_BM;
  p.n_tiles = (a.n + BN - 1) / BN;
  const int total_kb = (a.k + GEMM_BK - 1) / GEMM_BK;
  if (splits > total_kb) splits = total_kb;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  p.splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  p.reduce_add = (a.accumulate || p.splits > 1) ? 1 : 0;
  p.out_bytes = GEMM_BM * PN * (a.d_is_f32 ? 4 : 2);
  if (p.out_bytes < GEMM_BOX_BYTES) p.out_bytes = GEMM_BOX_BYTES;
  p.x_bytes = xptr != nullptr ? (GEMM_BM * PN * 2 < GEMM_BOX_BYTES ? GEMM_BOX_BYTES : GEMM_BM * PN * 2) : 0;
  const int fixed = 1024 /*alignment slack*/ + p.out_bytes + p.x_bytes + 2 * BN * 4 + 256 /*barriers*/;
  int stages = (GEMM_SMEM_LIMIT - fixed) / STAGE_BYTES;
  if (stages > GEMM_MAX_STAGES) stages = GEMM_MAX_STAGES;
  if (stages < 2) {
    vb_set_last_error("vb_gemm_bf16", "tile configuration does not fit shared memory");
    return VB_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  static const int debug_mode = getenv("VB_GEMM_DEBUG") ? atoi(getenv("VB_GEMM_DEBUG")) : 0;
  static const int debug_stages = getenv("VB_GEMM_STAGES") ? atoi(getenv("VB_GEMM_STAGES")) : 0;
  p.debug_mode = debug_mode;
  if (debug_stages >= 2 && debug_stages < stages) p.stages = stages = debug_stages;
  const int smem_bytes = fixed + stages * STAGE_BYTES;

  static int attr_bytes = 0;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN>;
  if (attr_bytes < smem_bytes) {
    VB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
    attr_bytes = GEMM_SMEM_LIMIT;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.splits;
  int grid = num_sms();
  if (a.max_ctas > 0 && a.max_ctas < grid) grid = a.max_ctas;
  if (tiles < grid) grid = tiles;
  kern<<<grid, GEMM_THREADS, smem_bytes, stream>>>(map_a, map_b, map_d, map_x, p);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

template <int BN>
static int dispatch_major(const vb_gemm_args& a, int splits, cudaStream_t s) {
  if (a.a_mn_major) {
    if (a.b_mn_major) return launch_gemm<BN, true, true>(a, splits, s);
    return launch_gemm<BN, true, false>(a, splits, s);
  }
  if (a.b_mn_major) return launch_gemm<BN, false, true>(a, splits, s);
  return launch_gemm<BN, false, false>(a, splits, s);
}

// tile-shape heuristic: the widest BN whose wave efficiency on the persistent grid is (nearly) the best
static void pick_config(const vb_gemm_args& a, int* bn_out, int* splits_out) {
  const int sms = (a.max_ctas > 0 && a.max_ctas < num_sms()) ? a.max_ctas : num_sms();
  const int m_tiles = (a.m + GEMM_BM - 1) / GEMM_BM;
  const int total_kb = (a.k + GEMM_BK - 1) / GEMM_BK;
  const bool can_split = a.d_is_f32 && a.accumulate;
  double best_cost = 1e30;
  int best_bn = 128, best_splits = 1;
  // measured on B200 at the ViLBERT shapes: a UTCHMMA has a ~110-cycle floor, so 128-wide tiles lose little against
  // 256-wide ones while doubling the number of tiles (these GEMMs are short of CTAs, not of MMA rate)
  const int bns[3] = {128, 64, 256};
  for (int bi = 0; bi < 3; ++bi) {
    const int bn = bns[bi];
    if (a.block_n != 0 && a.block_n != bn) continue;
    if (bn > 64 && a.n <= bn / 2) continue;  // do not waste most of a tile
    const int n_tiles = (a.n + bn - 1) / bn;
    const int max_s = can_split ? 16 : 1;
    for (int s = 1; s <= max_s; s *= 2) {
      if (a.splits != 0 && a.splits != s) continue;
      if (s > 1 && total_kb / s < 4) break;
      const long tiles = static_cast<long>(m_tiles) * n_tiles * s;
      const long waves = (tiles + sms - 1) / sms;
      const int kb = (total_kb + s - 1) / s;
      // per-tile time in SM cycles, fitted to tools/gemm_debug.py runs: a k-block (4 UTCHMMA) costs ~430 / 460 / 590
      // cycles at BN = 64 / 128 / 256, the epilogue ~5 cycles per column plus a fixed part, TMA reduce-adds a bit more
      const double kb_cost = bn == 64 ? 430.0 : (bn == 128 ? 460.0 : 590.0);
      const double tile_cost = kb * kb_cost + 400.0 + 5.0 * bn + (s > 1 ? 3.0 * bn : 0.0);
      const double cost = waves * tile_cost;
      if (cost < best_cost * 0.97) { best_cost = cost; best_bn = bn; best_splits = s; }
    }
  }
  *bn_out = best_bn;
  *splits_out = best_splits;
}

}  // namespace vb

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
  VB_REQUIRE(a.block_n == 0 || a.block_n == 64 || a.block_n == 128 || a.block_n == 256, "block_n must be 0, 64, 128 or 256");
  VB_REQUIRE(a.splits >= 0 && (a.splits <= 1 || (a.d_is_f32 && a.accumulate)), "split-K needs an fp32 accumulating output");
  VB_REQUIRE(!(a.accumulate && !a.d_is_f32), "accumulate needs an fp32 output");
  VB_REQUIRE(!(a.d_is_f32 && a.d_preact != nullptr), "d_preact only with a bf16 output");
  VB_REQUIRE(!(a.d_preact != nullptr && a.aux_mode != VB_AUX_NONE), "d_preact and aux share one staging panel: use one of them");
  VB_REQUIRE(!(a.d_is_f32 && a.aux_mode != VB_AUX_NONE), "aux only with a bf16 output");
  int bn = 128, splits = 1;
  pick_config(a, &bn, &splits);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (bn) {
    case 64: return dispatch_major<64>(a, splits, s);
    case 128: return dispatch_major<128>(a, splits, s);
    default: return dispatch_major<256>(a, splits, s);
  }
}
