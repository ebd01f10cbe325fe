// Fused softmax attention on tcgen05 / TMEM for the short sequences of ViLBERT (<= 128 queries, <= 128 keys per
// (sample, head); head width 64 or 128).  One CTA per (sample, head), 256 threads: two threads per query row (= TMEM lane),
// each owning 64 of the 128 key columns, so the softmax needs one exchange of (max, sum) per row through shared memory and the
// per-thread register tile is half a row (the 128-wide per-thread row of the first version was the critical path).
//
//   forward   S = Q K^T (UMMA 128x128xD) -> softmax(S*scale + mask) [+ dropout] -> P (bf16, swizzled smem) -> O = P V
//   backward  recompute S and dP' = dO V^T -> P, dS in smem -> dV = P^T dO, dK = dS^T Q, dQ = dS K  (5 UMMAs total)
//
// The same kernel serves self-attention (reference models/vilbert_facebook_arch.py:126-144) and both directions of
// the co-attention exchange (:253-294): Q, K and V are independent strided views, the additive key mask belongs to
// the key side.  Operand tiles are fetched by 3-D TMA ([batch, position, column], rows past the sequence end are
// zero-filled), so ragged region counts (36, 100) need no padding in HBM.
#include "common.cuh"
#include "../../include/vilbert_b200.h"
#include "tensormap.h"

namespace vb {

constexpr int ATT_CHUNK = 128 * 128;  // bytes of one [128 rows][64 bf16] 128B-swizzled tile

struct AttnKernelParams {
  float* lse;
  const float* mask_bias;
  __nv_bfloat16* out;
  __nv_bfloat16 *dq, *dk, *dv;
  long long ldo, lddq, lddk, lddv;
  long long q_rows, k_rows, bias_ld;   // rows between consecutive samples in the q- / k-side views; mask row stride
  const float* delta;                   // backward: external sum_k P dP per row ([batch, heads, 128]) or null
  int sq, sk, heads;
  float scale, p_drop;
  uint32_t site;
  const unsigned long long* seed;
};

constexpr int ATT_THREADS = 256;   // two threads per query row: thread (row, ch) owns key columns [64*ch, 64*ch + 64)

// 64 accumulator columns of this thread's row (2 x tcgen05.ld.32x32b.x32, one wait)
__device__ __forceinline__ void tmem_ld_row64(uint32_t taddr, float (&s)[64]) {
  uint32_t r0[32], r1[32];
  tmem_ld_32x32(taddr, r0);
  tmem_ld_32x32(taddr + 32, r1);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(r0[i]); s[32 + i] = __uint_as_float(r1[i]); }
}

// write 64 fp32 values of this thread's half row as bf16 into one [128][64] 128B-swizzled tile
__device__ __forceinline__ void store_half_row_bf16(uint8_t* tile, int row, const float (&v)[64]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 u;
    u.x = pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]); u.y = pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]);
    u.z = pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]); u.w = pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]);
    *reinterpret_cast<uint4*>(tile + swz128(row, j)) = u;
  }
}

// NC32 x 32 accumulator columns of this thread's row -> bf16 global
template <int NC32>
__device__ __forceinline__ void store_tmem_cols(uint32_t taddr, __nv_bfloat16* dst, bool ok) {
#pragma unroll
  for (int c = 0; c < NC32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + c * 32, r);
    tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        u.y = pack_bf16x2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        u.z = pack_bf16x2(__uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
        u.w = pack_bf16x2(__uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
        *reinterpret_cast<uint4*>(dst + c * 32 + i) = u;
      }
    }
  }
}

// 64-bit keep mask of half a query row: words 2*ch and 2*ch+1 of the row's 128-bit mask (bit k = key k is kept); the Philox
// counters are those of the whole-row mask, so forward and backward (and both halves) agree
__device__ __forceinline__ void attn_keep_half(uint64_t seed, uint32_t site, uint32_t row_index, uint32_t thr, int ch,
                                               uint32_t (&keep)[2]) {
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) m |= dropout_keep8(seed, site, row_index * 16u + (2 * ch + w) * 4 + q, thr) << (q * 8);
    keep[w] = m;
  }
}

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, D == 64 ? 2 : 1)   // 64-wide heads: 84 KB smem, 256 TMEM columns, <= 128 registers -> two CTAs per SM
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const AttnKernelParams p) {
  constexpr int NC = D / 64;
  constexpr uint32_t TMEM_COLS = 256;  // S: [0,128)  O: [128,128+D)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NC * ATT_CHUNK;
  uint8_t* sV = sK + NC * ATT_CHUNK;
  uint8_t* sP = sV + NC * ATT_CHUNK;
  float* sBias = reinterpret_cast<float*>(sP + 2 * ATT_CHUNK);
  float2* sRed = reinterpret_cast<float2*>(sBias + 128);          // [256] (max, sum) of every half row
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + ATT_THREADS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, ch = tid >> 7;
  const int h = blockIdx.x, b = blockIdx.y;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  if (tid < 128) sBias[tid] = tid < p.sk ? (p.mask_bias ? p.mask_bias[(long long)b * p.bias_ld + tid] : 0.f) : -INFINITY;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 3 * NC * ATT_CHUNK);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      tma_load_3d(sQ + c * ATT_CHUNK, &tm_q, &bars[0], h * D + c * 64, 0, b);
      tma_load_3d(sK + c * ATT_CHUNK, &tm_k, &bars[0], h * D + c * 64, 0, b);
      tma_load_3d(sV + c * ATT_CHUNK, &tm_v, &bars[0], h * D + c * 64, 0, b);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, false, false);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem, umma_smem_desc(smem_u32(sQ + c * ATT_CHUNK) + kk * 32, 16, 1024),
                  umma_smem_desc(smem_u32(sK + c * ATT_CHUNK) + kk * 32, 16, 1024), IDESC_S, (c | kk) ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  __syncwarp();
  tc_fence_after();

  // thread (row, ch): TMEM lane = row (sub-partition warp & 3), columns [64 ch, 64 ch + 64)
  const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float s[64];
  tmem_ld_row64(trow + ch * 64, s);
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 64; ++i) { s[i] = s[i] * p.scale + sBias[ch * 64 + i]; mx = fmaxf(mx, s[i]); }
  const float mx_safe = mx == -INFINITY ? 0.f : mx;   // a half row with no valid key (sk <= 64): every term is exp(-inf) = 0
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) { s[i] = __expf(s[i] - mx_safe); sum += s[i]; }
  sRed[tid] = make_float2(mx, sum);
  __syncthreads();
  const float2 other = sRed[tid ^ 128];
  const float m_all = fmaxf(mx, other.x);                               // finite: key 0 is always valid or masked with -10000
  const float mine = mx == -INFINITY ? 0.f : __expf(mx - m_all), theirs = other.x == -INFINITY ? 0.f : __expf(other.x - m_all);
  const float total = sum * mine + other.y * theirs;
  const float inv = mine / total;
  if (ch == 0) p.lse[((long long)b * p.heads + h) * 128 + row] = m_all + logf(total);
  if (p.p_drop > 0.f && p.seed) {
    uint32_t keep[2];
    attn_keep_half(*p.seed, p.site, (uint32_t)((b * p.heads + h) * 128 + row), dropout_threshold(p.p_drop), ch, keep);
    const float invk = inv / (1.f - p.p_drop);
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] = ((keep[i >> 5] >> (i & 31)) & 1u) ? s[i] * invk : 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] *= inv;
  }
  store_half_row_bf16(sP + ch * ATT_CHUNK, row, s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t IDESC_O = umma_idesc_bf16(128, D, false, true);
    const int nkk = (p.sk + 15) >> 4;
    for (int kk = 0; kk < nkk; ++kk)
      umma_bf16(tmem + 128, umma_smem_desc(smem_u32(sP + (kk >> 2) * ATT_CHUNK) + (kk & 3) * 32, 16, 1024),
                umma_smem_desc(smem_u32(sV) + kk * 2048, ATT_CHUNK, 1024), IDESC_O, kk ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 1);
  __syncwarp();
  tc_fence_after();
  store_tmem_cols<D / 64>(trow + 128 + ch * (D / 2), p.out + ((long long)b * p.q_rows + row) * p.ldo + h * D + ch * (D / 2), row < p.sq);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, TMEM_COLS); }
}

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const AttnKernelParams p) {
  constexpr int NC = D / 64;
  constexpr uint32_t TMEM_COLS = 512;  // S/dQ: [0,128)  dP: [128,256)  dV: [256,256+D)  dK: [384,384+D)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NC * ATT_CHUNK;
  uint8_t* sV = sK + NC * ATT_CHUNK;
  uint8_t* sdO = sV + NC * ATT_CHUNK;
  uint8_t* sP = sdO + NC * ATT_CHUNK;
  uint8_t* sdS = sP + 2 * ATT_CHUNK;
  float* sBias = reinterpret_cast<float*>(sdS + 2 * ATT_CHUNK);
  float2* sRed = reinterpret_cast<float2*>(sBias + 128);          // [256]: .x = partial sum_k P dP of every half row
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + ATT_THREADS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, ch = tid >> 7;
  const int h = blockIdx.x, b = blockIdx.y;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  if (tid < 128) sBias[tid] = tid < p.sk ? (p.mask_bias ? p.mask_bias[(long long)b * p.bias_ld + tid] : 0.f) : -INFINITY;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 4 * NC * ATT_CHUNK);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      tma_load_3d(sQ + c * ATT_CHUNK, &tm_q, &bars[0], h * D + c * 64, 0, b);
      tma_load_3d(sK + c * ATT_CHUNK, &tm_k, &bars[0], h * D + c * 64, 0, b);
      tma_load_3d(sV + c * ATT_CHUNK, &tm_v, &bars[0], h * D + c * 64, 0, b);
      tma_load_3d(sdO + c * ATT_CHUNK, &tm_do, &bars[0], h * D + c * 64, 0, b);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, false, false);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem, umma_smem_desc(smem_u32(sQ + c * ATT_CHUNK) + kk * 32, 16, 1024),
                  umma_smem_desc(smem_u32(sK + c * ATT_CHUNK) + kk * 32, 16, 1024), IDESC_S, (c | kk) ? 1u : 0u);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem + 128, umma_smem_desc(smem_u32(sdO + c * ATT_CHUNK) + kk * 32, 16, 1024),
                  umma_smem_desc(smem_u32(sV + c * ATT_CHUNK) + kk * 32, 16, 1024), IDESC_S, (c | kk) ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  __syncwarp();
  tc_fence_after();

  const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const bool drop = p.p_drop > 0.f && p.seed;
  uint32_t keep[2] = {0xffffffffu, 0xffffffffu};
  if (drop) attn_keep_half(*p.seed, p.site, (uint32_t)((b * p.heads + h) * 128 + row), dropout_threshold(p.p_drop), ch, keep);
  const float invk = drop ? 1.f / (1.f - p.p_drop) : 1.f;
  const float lse = p.lse[((long long)b * p.heads + h) * 128 + row];

  float pr[64];
  tmem_ld_row64(trow + ch * 64, pr);
#pragma unroll
  for (int i = 0; i < 64; ++i) pr[i] = __expf(pr[i] * p.scale + sBias[ch * 64 + i] - lse);
  {
    // P after dropout feeds dV = P_drop^T dO
    float pd[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) pd[i] = ((keep[i >> 5] >> (i & 31)) & 1u) ? pr[i] * invk : 0.f;
    store_half_row_bf16(sP + ch * ATT_CHUNK, row, pd);
  }
  float dp[64];
  tmem_ld_row64(trow + 128 + ch * 64, dp);
  float part = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    dp[i] = ((keep[i >> 5] >> (i & 31)) & 1u) ? dp[i] * invk : 0.f;
    part += pr[i] * dp[i];
  }
  sRed[tid] = make_float2(part, 0.f);
  __syncthreads();
  // key-blocked calls (sequences above 128) pass the row's sum over ALL key blocks; a single block computes it here
  const float drow = p.delta ? p.delta[((long long)b * p.heads + h) * 128 + row] : part + sRed[tid ^ 128].x;
#pragma unroll
  for (int i = 0; i < 64; ++i) pr[i] = pr[i] * (dp[i] - drow) * p.scale;   // dS (scaled so that dQ = dS K, dK = dS^T Q)
  store_half_row_bf16(sdS + ch * ATT_CHUNK, row, pr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t IDESC_T = umma_idesc_bf16(128, D, true, true);   // A^T (MN-major) x MN-major B
    constexpr uint32_t IDESC_Q = umma_idesc_bf16(128, D, false, true);
    const int nq = (p.sq + 15) >> 4, nk = (p.sk + 15) >> 4;
    for (int kk = 0; kk < nq; ++kk)  // dV[key, d] = sum_q Pd[q, key] dO[q, d]
      umma_bf16(tmem + 256, umma_smem_desc(smem_u32(sP) + kk * 2048, ATT_CHUNK, 1024),
                umma_smem_desc(smem_u32(sdO) + kk * 2048, ATT_CHUNK, 1024), IDESC_T, kk ? 1u : 0u);
    for (int kk = 0; kk < nq; ++kk)  // dK[key, d] = sum_q dS[q, key] Q[q, d]
      umma_bf16(tmem + 384, umma_smem_desc(smem_u32(sdS) + kk * 2048, ATT_CHUNK, 1024),
                umma_smem_desc(smem_u32(sQ) + kk * 2048, ATT_CHUNK, 1024), IDESC_T, kk ? 1u : 0u);
    for (int kk = 0; kk < nk; ++kk)  // dQ[q, d] = sum_key dS[q, key] K[key, d]
      umma_bf16(tmem, umma_smem_desc(smem_u32(sdS + (kk >> 2) * ATT_CHUNK) + (kk & 3) * 32, 16, 1024),
                umma_smem_desc(smem_u32(sK) + kk * 2048, ATT_CHUNK, 1024), IDESC_Q, kk ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 1);
  __syncwarp();
  tc_fence_after();
  constexpr int HD = D / 2;   // each of the two threads of a row stores half of the head width
  store_tmem_cols<D / 64>(trow + ch * HD, p.dq + ((long long)b * p.q_rows + row) * p.lddq + h * D + ch * HD, row < p.sq);
  store_tmem_cols<D / 64>(trow + 256 + ch * HD, p.dv + ((long long)b * p.k_rows + row) * p.lddv + h * D + ch * HD, row < p.sk);
  store_tmem_cols<D / 64>(trow + 384 + ch * HD, p.dk + ((long long)b * p.k_rows + row) * p.lddk + h * D + ch * HD, row < p.sk);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, TMEM_COLS); }
}

template <int D>
static constexpr int attn_smem_bytes(bool bwd) {
  return (bwd ? 4 : 3) * (D / 64) * ATT_CHUNK + (bwd ? 4 : 2) * ATT_CHUNK + 128 * 4 + ATT_THREADS * 8 + 64 + 1024;
}

static int make_view_map(CUtensorMap* m, const void* base, int width, int seq, int batch, long long ld, long long batch_rows) {
  // [batch, seq, width] view with row stride ld and batch_rows rows between samples (> seq when the view is one block of a
  // longer sequence: rows past `seq` are zero-filled by TMA either way); box = 64 columns x 128 rows x 1 sample
  return make_tensor_map_3d(m, base, width, seq, batch, ld, batch_rows * ld, 64, 128, 1);
}

template <int D>
static int launch_attn(const vb_attn_args& a, bool bwd, cudaStream_t stream) {
  CUtensorMap mq, mk, mv, mdo;
  const int width = a.heads * D;
  int rc;
  const long long q_rows = a.q_batch_rows > 0 ? a.q_batch_rows : a.sq, k_rows = a.k_batch_rows > 0 ? a.k_batch_rows : a.sk;
  if ((rc = make_view_map(&mq, a.q, width, a.sq, a.batch, a.ldq, q_rows)) != VB_OK) return rc;
  if ((rc = make_view_map(&mk, a.k, width, a.sk, a.batch, a.ldk, k_rows)) != VB_OK) return rc;
  if ((rc = make_view_map(&mv, a.v, width, a.sk, a.batch, a.ldv, k_rows)) != VB_OK) return rc;
  AttnKernelParams p;
  p.lse = a.lse; p.mask_bias = a.mask_bias; p.out = (__nv_bfloat16*)a.out;
  p.dq = (__nv_bfloat16*)a.dq; p.dk = (__nv_bfloat16*)a.dk; p.dv = (__nv_bfloat16*)a.dv;
  p.ldo = a.ldo; p.lddq = a.lddq; p.lddk = a.lddk; p.lddv = a.lddv;
  p.q_rows = q_rows; p.k_rows = k_rows; p.bias_ld = a.bias_ld > 0 ? a.bias_ld : a.sk; p.delta = a.delta;
  p.sq = a.sq; p.sk = a.sk; p.heads = a.heads; p.scale = a.scale; p.p_drop = a.p_drop; p.site = a.site;
  p.seed = (const unsigned long long*)a.seed;
  dim3 grid(a.heads, a.batch);
  if (!bwd) {
    static bool attr = false;
    if (!attr) {
      VB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<D>(false)));
      attr = true;
    }
    attn_fwd_kernel<D><<<grid, ATT_THREADS, attn_smem_bytes<D>(false), stream>>>(mq, mk, mv, p);
  } else {
    if ((rc = make_view_map(&mdo, a.dout, width, a.sq, a.batch, a.lddo, q_rows)) != VB_OK) return rc;
    static bool attr = false;
    if (!attr) {
      VB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<D>(true)));
      attr = true;
    }
    attn_bwd_kernel<D><<<grid, ATT_THREADS, attn_smem_bytes<D>(true), stream>>>(mq, mk, mv, mdo, p);
  }
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

static int check_attn(const vb_attn_args* a, bool bwd) {
  VB_REQUIRE(a != nullptr, "null args");
  VB_REQUIRE(a->q && a->k && a->v && a->lse, "q, k, v and lse are required");
  VB_REQUIRE(a->d == 64 || a->d == 128, "head width must be 64 or 128");
  VB_REQUIRE(a->sq >= 1 && a->sq <= 128 && a->sk >= 1 && a->sk <= 128, "1 <= sq, sk <= 128");
  VB_REQUIRE(a->batch >= 1 && a->heads >= 1, "batch and heads must be positive");
  VB_REQUIRE((a->q_batch_rows == 0 || a->q_batch_rows >= a->sq) && (a->k_batch_rows == 0 || a->k_batch_rows >= a->sk) &&
             (a->bias_ld == 0 || a->bias_ld >= a->sk), "per-sample row counts must cover the block");
  VB_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0, "leading dimensions must be multiples of 8");
  VB_REQUIRE(((uintptr_t)a->q & 15) == 0 && ((uintptr_t)a->k & 15) == 0 && ((uintptr_t)a->v & 15) == 0, "q/k/v must be 16-byte aligned");
  VB_REQUIRE(a->p_drop >= 0.f && a->p_drop < 1.f && (a->p_drop == 0.f || a->seed), "dropout p in [0,1) with a device seed");
  if (!bwd) {
    VB_REQUIRE(a->out && a->ldo % 8 == 0 && ((uintptr_t)a->out & 15) == 0, "out missing or misaligned");
  } else {
    VB_REQUIRE(a->dout && a->dq && a->dk && a->dv, "dout, dq, dk, dv are required");
    VB_REQUIRE(a->lddo % 8 == 0 && a->lddq % 8 == 0 && a->lddk % 8 == 0 && a->lddv % 8 == 0, "leading dimensions must be multiples of 8");
    VB_REQUIRE(((uintptr_t)a->dout & 15) == 0 && ((uintptr_t)a->dq & 15) == 0 && ((uintptr_t)a->dk & 15) == 0 && ((uintptr_t)a->dv & 15) == 0,
               "gradient pointers must be 16-byte aligned");
  }
  return VB_OK;
}

}  // namespace vb

extern "C" int vb_attention_fwd(const vb_attn_args* a, void* stream) {
  using namespace vb;
  int rc = check_attn(a, false);
  if (rc != VB_OK) return rc;
  return a->d == 64 ? launch_attn<64>(*a, false, (cudaStream_t)stream) : launch_attn<128>(*a, false, (cudaStream_t)stream);
}

extern "C" int vb_attention_bwd(const vb_attn_args* a, void* stream) {
  using namespace vb;
  int rc = check_attn(a, true);
  if (rc != VB_OK) return rc;
  return a->d == 64 ? launch_attn<64>(*a, true, (cudaStream_t)stream) : launch_attn<128>(*a, true, (cudaStream_t)stream);
}
