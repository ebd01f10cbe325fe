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
#include "common.cuh"
#include "../../include/vilbert_b200.h"
#include "tensormap.h"

namespace vb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;       // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 192; // 6 warps
constexpr int GEMM_ACC_STAGES = 2;

struct GemmKernelParams {
  void* d;
  void* d_preact;
  const float* scale;
  const float* bias;
  const __nv_bfloat16* aux;
  long long ldd, ld_preact, ld_aux;
  int m, n, k;
  int d_is_f32, accumulate, act, aux_mode;
  int splits, kb_per_split;
  int m_tiles, n_tiles;
};

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int BUDGET = 200 * 1024;
  static constexpr int STAGES = (BUDGET - BAR_BYTES) / STAGE_BYTES < 8 ? (BUDGET - BAR_BYTES) / STAGE_BYTES : 8;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024 /*alignment slack*/;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case VB_ACT_GELU: return gelu_erf(v);
    case VB_ACT_RELU: return fmaxf(v, 0.0f);
    case VB_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const GemmKernelParams p) {
  using S = GemmSmem<BN>;
  constexpr int STAGES = S::STAGES;
  constexpr uint32_t TMEM_COLS = GEMM_ACC_STAGES * BN;  // 128, 256 or 512: all powers of two >= 32
  constexpr uint32_t IDESC = umma_idesc_bf16(GEMM_BM, BN, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * S::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + GEMM_ACC_STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * GEMM_ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < GEMM_ACC_STAGES; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);  // one arrive per epilogue warp
    }
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
  const int num_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mn = tile % (p.m_tiles * p.n_tiles);
        const int split = tile / (p.m_tiles * p.n_tiles);
        const int m0 = (mn % p.m_tiles) * GEMM_BM;
        const int n0 = (mn / p.m_tiles) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * S::A_BYTES;
          uint8_t* sb = smem_b + stage * S::B_BYTES;
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
      const int split = tile / (p.m_tiles * p.n_tiles);
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(total_kb, kb0 + p.kb_per_split);
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem_a + stage * S::A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * S::B_BYTES);
#pragma unroll
          for (int kk = 0; kk < GEMM_BK / 16; ++kk) {
            uint64_t da, db;
            if constexpr (A_MN) da = umma_smem_desc(sa + kk * 2048, GEMM_BK * 128, 1024);
            else                da = umma_smem_desc(sa + kk * 32, 16, 1024);
            if constexpr (B_MN) db = umma_smem_desc(sb + kk * 2048, GEMM_BK * 128, 1024);
            else                db = umma_smem_desc(sb + kk * 32, 16, 1024);
            umma_bf16(tmem_d, da, db, IDESC, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (lane == 0) umma_commit(&tmem_full_bar[acc]);
      __syncwarp();
      if (++acc == GEMM_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM sub-partition this warp may read: lanes [32q, 32q+32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mn = tile % (p.m_tiles * p.n_tiles);
      const int m0 = (mn % p.m_tiles) * GEMM_BM;
      const int n0 = (mn / p.m_tiles) * BN;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.m;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row_ok && col0 < p.n) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          const int ncols = min(32, p.n - col0);  // multiple of 8 (host checks n % 8 == 0)
          if (p.scale != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (i < ncols) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(p.scale + col0 + i));
                v[i] *= s.x; v[i + 1] *= s.y; v[i + 2] *= s.z; v[i + 3] *= s.w;
              }
            }
          }
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (i < ncols) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
                v[i] += s.x; v[i + 1] += s.y; v[i + 2] += s.z; v[i + 3] += s.w;
              }
            }
          }
          if (p.d_preact != nullptr) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.d_preact) + static_cast<long long>(row) * p.ld_preact + col0;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              if (i < ncols) {
                uint4 o;
                o.x = pack_bf16x2(v[i], v[i + 1]); o.y = pack_bf16x2(v[i + 2], v[i + 3]);
                o.z = pack_bf16x2(v[i + 4], v[i + 5]); o.w = pack_bf16x2(v[i + 6], v[i + 7]);
                *reinterpret_cast<uint4*>(dst + i) = o;
              }
            }
          }
          if (p.aux_mode != VB_AUX_NONE) {
            const __nv_bfloat16* src = p.aux + static_cast<long long>(row) * p.ld_aux + col0;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              if (i < ncols) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(src + i));
                const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
                const float av[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
                if (p.aux_mode == VB_AUX_ADD) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[i + j] += av[j];
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[i + j] *= gelu_erf_grad(av[j]);
                }
              }
            }
          }
          if (p.act != VB_ACT_NONE) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = apply_act(v[i], p.act);
          }
          if (p.d_is_f32) {
            float* dst = reinterpret_cast<float*>(p.d) + static_cast<long long>(row) * p.ldd + col0;
            if (p.accumulate) {
              if (p.splits > 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < ncols) atomicAdd(dst + i, v[i]);
              } else {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  if (i < ncols) {
                    float4 o = *reinterpret_cast<float4*>(dst + i);
                    o.x += v[i]; o.y += v[i + 1]; o.z += v[i + 2]; o.w += v[i + 3];
                    *reinterpret_cast<float4*>(dst + i) = o;
                  }
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (i < ncols) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.d) + static_cast<long long>(row) * p.ldd + col0;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              if (i < ncols) {
                uint4 o;
                o.x = pack_bf16x2(v[i], v[i + 1]); o.y = pack_bf16x2(v[i + 2], v[i + 3]);
                o.z = pack_bf16x2(v[i + 4], v[i + 5]); o.w = pack_bf16x2(v[i + 6], v[i + 7]);
                *reinterpret_cast<uint4*>(dst + i) = o;
              }
            }
          }
        }
      }
      // accumulator stage drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == GEMM_ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
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
  using S = GemmSmem<BN>;
  CUtensorMap map_a, map_b;
  int rc;
  // K-major operand: global [rows, K] -> box {64 (k), rows_per_tile}; MN-major: global [K, rows] -> box {64 (mn), 64 (k)}
  if (A_MN) rc = make_tensor_map_2d(&map_a, a.a, /*inner*/ a.m, /*outer*/ a.k, a.lda, 64, GEMM_BK);
  else      rc = make_tensor_map_2d(&map_a, a.a, a.k, a.m, a.lda, GEMM_BK, GEMM_BM);
  if (rc != VB_OK) return rc;
  if (B_MN) rc = make_tensor_map_2d(&map_b, a.b, a.n, a.k, a.ldb, 64, GEMM_BK);
  else      rc = make_tensor_map_2d(&map_b, a.b, a.k, a.n, a.ldb, GEMM_BK, BN);
  if (rc != VB_OK) return rc;

  GemmKernelParams p;
  p.d = a.d; p.d_preact = a.d_preact; p.scale = a.scale; p.bias = a.bias;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a.aux);
  p.ldd = a.ldd; p.ld_preact = a.ld_preact; p.ld_aux = a.ld_aux;
  p.m = a.m; p.n = a.n; p.k = a.k;
  p.d_is_f32 = a.d_is_f32; p.accumulate = a.accumulate; p.act = a.act; p.aux_mode = a.aux_mode;
  p.m_tiles = (a.m + GEMM_BM - 1) / GEMM_BM;
  p.n_tiles = (a.n + BN - 1) / BN;
  const int total_kb = (a.k + GEMM_BK - 1) / GEMM_BK;
  if (splits > total_kb) splits = total_kb;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  p.splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split

  static bool attr_set = false;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN>;
  if (!attr_set) {
    VB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.splits;
  int grid = num_sms();
  if (a.max_ctas > 0 && a.max_ctas < grid) grid = a.max_ctas;
  if (tiles < grid) grid = tiles;
  kern<<<grid, GEMM_THREADS, S::TOTAL, stream>>>(map_a, map_b, p);
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
  const int bns[3] = {256, 128, 64};
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
      // per-tile time ~ main loop (bn/64 units per k-block) + fixed epilogue/fill cost (in the same units)
      const double tile_cost = kb * (bn / 64.0) + 6.0 + bn / 32.0 + (s > 1 ? bn / 16.0 : 0.0);
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
  int bn = 128, splits = 1;
  pick_config(a, &bn, &splits);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (bn) {
    case 64: return dispatch_major<64>(a, splits, s);
    case 128: return dispatch_major<128>(a, splits, s);
    default: return dispatch_major<256>(a, splits, s);
  }
}
