// Bandwidth-bound row-wise kernels of the ViLBERT hot path (coalesced 16-byte accesses, one warp per row,
// warp-shuffle reductions, fp32 math on bf16 storage):
//   LayerNorm(+dropout +residual) forward / backward      reference models/vilbert_facebook_arch.py:63-76,156-160,197-201
//   text embeddings gather + LayerNorm forward / backward  transformers BertEmbeddings (called at :524)
//   column sums (bias gradients), casts, additive masks, element-wise dropout / activation backward
#include <type_traits>

#include "common.cuh"
#include "../../include/vilbert_b200.h"

namespace vb {

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void ld8f(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// keep decisions of 8 consecutive elements starting at element index e (e % 8 == 0) of dropout site `site`
__device__ __forceinline__ void keep8(uint64_t seed, uint32_t site, uint32_t e, uint32_t thr, bool (&k)[8]) {
  const uint32_t m = dropout_keep8(seed, site, e >> 3, thr);
#pragma unroll
  for (int i = 0; i < 8; ++i) k[i] = (m >> i) & 1u;
}

constexpr int LN_WARPS = 8;

struct LnParams {
  const __nv_bfloat16* x;     // [M,H] dense output (pre-dropout)
  const __nv_bfloat16* res;   // [M,H] residual or null
  const float* res32;         // the same residual in fp32 (takes precedence): the residual stream is carried in fp32
  float* y32;                 // optional fp32 copy of the output (the next block's residual)
  long long ldres32, ldy32;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* y;           // fwd out
  float* mean;                // [M]
  float* rstd;                // [M]
  // backward
  const __nv_bfloat16* dy;    // [M,H]
  __nv_bfloat16* dx;          // grad wrt x (after the input-dropout mask); may be null
  __nv_bfloat16* dres;        // grad wrt res (= grad wrt the pre-norm sum); may be null
  float* dgamma;              // atomically accumulated
  float* dbeta;
  float* dbias;               // optional: column sum of dx (bias gradient of the dense that produced x)
  long long ldx, ldres, ldy, lddy, lddx, lddres;
  int m, h;
  float eps;
  float p_in, p_out;          // dropout on x before the residual add / on the normalised output
  uint32_t site_in, site_out;
  const unsigned long long* seed;  // device pointer (null = no dropout)
};

template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint64_t seed = p.seed ? *p.seed : 0ull;
  const bool drop_in = p.p_in > 0.f && p.seed, drop_out = p.p_out > 0.f && p.seed;
  const uint32_t thr_in = dropout_threshold(p.p_in), thr_out = dropout_threshold(p.p_out);
  const float inv_in = drop_in ? 1.f / (1.f - p.p_in) : 1.f, inv_out = drop_out ? 1.f / (1.f - p.p_out) : 1.f;
  for (int row = blockIdx.x * LN_WARPS + warp; row < p.m; row += gridDim.x * LN_WARPS) {
    float v[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (lane + 32 * j) * 8;
      ld8(p.x + (long long)row * p.ldx + col, v[j]);
      if (drop_in) {
        bool k[8];
        keep8(seed, p.site_in, (uint32_t)row * (uint32_t)p.h + col, thr_in, k);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j][i] = k[i] ? v[j][i] * inv_in : 0.f;
      }
      if (p.res32) {
        float r[8];
        ld8f(p.res32 + (long long)row * p.ldres32 + col, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j][i] += r[i];
      } else if (p.res) {
        float r[8];
        ld8(p.res + (long long)row * p.ldres + col, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j][i] += r[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += v[j][i];
    }
    const float mean = warp_sum(sum) / (float)p.h;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; sq += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)p.h + p.eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (lane + 32 * j) * 8;
      float g[8], b[8], o[8];
      ld8f(p.gamma + col, g);
      ld8f(p.beta + col, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (v[j][i] - mean) * rstd * g[i] + b[i];
      if (drop_out) {
        bool k[8];
        keep8(seed, p.site_out, (uint32_t)row * (uint32_t)p.h + col, thr_out, k);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = k[i] ? o[i] * inv_out : 0.f;
      }
      st8(p.y + (long long)row * p.ldy + col, o);
      if (p.y32) {
        float* o32 = p.y32 + (long long)row * p.ldy32 + col;
        *reinterpret_cast<float4*>(o32) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(o32 + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
    if (lane == 0) {
      if (p.mean) p.mean[row] = mean;
      if (p.rstd) p.rstd[row] = rstd;
    }
  }
}

// Column partials (dgamma / dbeta / dbias): every warp parks its per-lane sums in its own smem row (plain stores), the
// block adds the LN_WARPS rows per column and issues ONE global atomicAdd per column (destination zeroed by the caller
// at the start of every backward pass).
template <int NV>
__device__ __forceinline__ void flush_colsum(const float (&acc)[NV][8], float* smem, float* gdst, int h, int lane, int warp) {
  __syncthreads();   // previous use of `smem` is over
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float* dst = smem + warp * h + (lane + 32 * j) * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < h; c += blockDim.x) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) sum += smem[w * h + c];
    atomicAdd(&gdst[c], sum);
  }
}

template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const LnParams p) {
  extern __shared__ float ln_smem[];   // [LN_WARPS][h]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint64_t seed = p.seed ? *p.seed : 0ull;
  const bool drop_in = p.p_in > 0.f && p.seed, drop_out = p.p_out > 0.f && p.seed;
  const uint32_t thr_in = dropout_threshold(p.p_in), thr_out = dropout_threshold(p.p_out);
  const float inv_in = drop_in ? 1.f / (1.f - p.p_in) : 1.f, inv_out = drop_out ? 1.f / (1.f - p.p_out) : 1.f;
  const float inv_h = 1.0f / (float)p.h;
  float dg[NV][8], db[NV][8], dbx[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) { dg[j][i] = 0.f; db[j][i] = 0.f; dbx[j][i] = 0.f; }
  for (int row = blockIdx.x * LN_WARPS + warp; row < p.m; row += gridDim.x * LN_WARPS) {
    // issue every global load of the row first (3 x NV independent 16-byte loads per lane)
    uint4 xr[NV], rr[NV], dr[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (lane + 32 * j) * 8;
      xr[j] = *reinterpret_cast<const uint4*>(p.x + (long long)row * p.ldx + col);
      dr[j] = *reinterpret_cast<const uint4*>(p.dy + (long long)row * p.lddy + col);
      if (p.res && !p.res32) rr[j] = *reinterpret_cast<const uint4*>(p.res + (long long)row * p.ldres + col);
    }
    const float mean = p.mean[row], rstd = p.rstd[row];
    float xh[NV][8], dxh[NV][8];
    uint32_t keep_in[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (lane + 32 * j) * 8;
      float v[8], d[8], g[8];
      { const float2 a = unpack_bf16x2(xr[j].x), b = unpack_bf16x2(xr[j].y), c = unpack_bf16x2(xr[j].z), e = unpack_bf16x2(xr[j].w);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = e.x; v[7] = e.y; }
      { const float2 a = unpack_bf16x2(dr[j].x), b = unpack_bf16x2(dr[j].y), c = unpack_bf16x2(dr[j].z), e = unpack_bf16x2(dr[j].w);
        d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y; d[4] = c.x; d[5] = c.y; d[6] = e.x; d[7] = e.y; }
      keep_in[j] = 0xFFu;
      if (drop_in) {
        bool k[8];
        keep8(seed, p.site_in, (uint32_t)row * (uint32_t)p.h + col, thr_in, k);
        keep_in[j] = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] = k[i] ? v[i] * inv_in : 0.f; keep_in[j] |= (k[i] ? 1u : 0u) << i; }
      }
      if (p.res32) {
        float r[8];
        ld8f(p.res32 + (long long)row * p.ldres32 + col, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += r[i];
      } else if (p.res) {
        const float2 a = unpack_bf16x2(rr[j].x), b = unpack_bf16x2(rr[j].y), c = unpack_bf16x2(rr[j].z), e = unpack_bf16x2(rr[j].w);
        v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += e.x; v[7] += e.y;
      }
      ld8f(p.gamma + col, g);
      if (drop_out) {
        bool k[8];
        keep8(seed, p.site_out, (uint32_t)row * (uint32_t)p.h + col, thr_out, k);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = k[i] ? d[i] * inv_out : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[j][i] = (v[i] - mean) * rstd;
        dxh[j][i] = d[i] * g[i];
        dg[j][i] += d[i] * xh[j][i];
        db[j][i] += d[i];
        s1 += dxh[j][i];
        s2 += dxh[j][i] * xh[j][i];
      }
    }
    const float c1 = warp_sum(s1) * inv_h, c2 = warp_sum(s2) * inv_h;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (lane + 32 * j) * 8;
      float ds[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ds[i] = rstd * (dxh[j][i] - c1 - xh[j][i] * c2);
      if (p.dres) st8(p.dres + (long long)row * p.lddres + col, ds);
      if (drop_in) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ds[i] = ((keep_in[j] >> i) & 1u) ? ds[i] * inv_in : 0.f;
      }
      if (p.dx) st8(p.dx + (long long)row * p.lddx + col, ds);
#pragma unroll
      for (int i = 0; i < 8; ++i) dbx[j][i] += ds[i];
    }
  }
  if (p.dgamma) flush_colsum<NV>(dg, ln_smem, p.dgamma, p.h, lane, warp);
  if (p.dbeta) flush_colsum<NV>(db, ln_smem, p.dbeta, p.h, lane, warp);
  if (p.dbias) flush_colsum<NV>(dbx, ln_smem, p.dbias, p.h, lane, warp);
}

static int ln_grid(int m, int rows_per_warp) {
  int g = (m + LN_WARPS * rows_per_warp - 1) / (LN_WARPS * rows_per_warp);
  return g < 1 ? 1 : g;
}

template <typename F>
static int dispatch_nv(int h, F&& f) {
  switch (h / 256) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
    case 4: return f(std::integral_constant<int, 4>());
    case 8: return f(std::integral_constant<int, 8>());
    default: vb_set_last_error("hidden size", "supported: 256, 512, 768, 1024, 2048"); return VB_ERR_UNSUPPORTED;
  }
}

// ------------------------------------------------------------------------------------------------ text embeddings
struct EmbParams {
  const int* ids;        // [B*T]
  const int* type_ids;   // [B*T] or null (= all zero)
  const float* word;     // [V,H] fp32 master tables
  const float* pos;      // [P,H]
  const float* type;     // [2,H]
  const float* gamma;
  const float* beta;
  __nv_bfloat16* y;      // [B*T,H]
  float* y32;            // optional fp32 copy (residual stream)
  float* mean;
  float* rstd;
  const __nv_bfloat16* dy;
  float* dword; float* dpos; float* dtype; float* dgamma; float* dbeta;   // fp32 gradients, atomically accumulated
  int b, t, h, vocab;
  float eps, p_out;
  uint32_t site_out;
  const unsigned long long* seed;
};

template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) emb_fwd_kernel(const EmbParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t seed = p.seed ? *p.seed : 0ull;
  const bool drop_out = p.p_out > 0.f && p.seed;
  const uint32_t thr_out = dropout_threshold(p.p_out);
  const float inv_out = drop_out ? 1.f / (1.f - p.p_out) : 1.f;
  const int m = p.b * p.t;
  for (int row = blockIdx.x * LN_WARPS + warp; row < m; row += gridDim.x * LN_WARPS) {
    const int id = p.ids[row], tt = p.type_ids ? p.type_ids[row] : 0, pos = row % p.t;
    float v[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (lane + 32 * j) * 8;
      float a[8], b[8], c[8];
      ld8f(p.word + (long long)id * p.h + col, a);
      ld8f(p.type + (long long)tt * p.h + col, b);
      ld8f(p.pos + (long long)pos * p.h + col, c);
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[j][i] = (a[i] + b[i]) + c[i]; sum += v[j][i]; }
    }
    const float mean = warp_sum(sum) / (float)p.h;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; sq += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)p.h + p.eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (lane + 32 * j) * 8;
      float g[8], b[8], o[8];
      ld8f(p.gamma + col, g);
      ld8f(p.beta + col, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (v[j][i] - mean) * rstd * g[i] + b[i];
      if (drop_out) {
        bool k[8];
        keep8(seed, p.site_out, (uint32_t)row * (uint32_t)p.h + col, thr_out, k);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = k[i] ? o[i] * inv_out : 0.f;
      }
      st8(p.y + (long long)row * p.h + col, o);
      if (p.y32) {
        float* o32 = p.y32 + (long long)row * p.h + col;
        *reinterpret_cast<float4*>(o32) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(o32 + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
    if (lane == 0) { p.mean[row] = mean; p.rstd[row] = rstd; }
  }
}

// One block per position t; its warps walk the batch.  Position gradients are therefore owned by one block (plain
// store after an in-block reduction); word rows use vector atomics (duplicates are rare; PAD id 0 is padding_idx ->
// no gradient); type / gamma / beta partials are reduced per block, then one atomic per column.
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) emb_bwd_kernel(const EmbParams p) {
  extern __shared__ float ln_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t seed = p.seed ? *p.seed : 0ull;
  const bool drop_out = p.p_out > 0.f && p.seed;
  const uint32_t thr_out = dropout_threshold(p.p_out);
  const float inv_out = drop_out ? 1.f / (1.f - p.p_out) : 1.f;
  const int pos = blockIdx.x;
  float dg[NV][8], db[NV][8], dp[NV][8], dt0[NV][8], dt1[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) { dg[j][i] = 0.f; db[j][i] = 0.f; dp[j][i] = 0.f; dt0[j][i] = 0.f; dt1[j][i] = 0.f; }
  for (int bi = warp; bi < p.b; bi += LN_WARPS) {
    const int row = bi * p.t + pos;
    const int id = p.ids[row], tt = p.type_ids ? p.type_ids[row] : 0;
    const float mean = p.mean[row], rstd = p.rstd[row];
    float xh[NV][8], dxh[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (lane + 32 * j) * 8;
      float a[8], b[8], c[8], d[8], g[8];
      ld8f(p.word + (long long)id * p.h + col, a);
      ld8f(p.type + (long long)tt * p.h + col, b);
      ld8f(p.pos + (long long)pos * p.h + col, c);
      ld8(p.dy + (long long)row * p.h + col, d);
      ld8f(p.gamma + col, g);
      if (drop_out) {
        bool k[8];
        keep8(seed, p.site_out, (uint32_t)row * (uint32_t)p.h + col, thr_out, k);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = k[i] ? d[i] * inv_out : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[j][i] = (((a[i] + b[i]) + c[i]) - mean) * rstd;
        dxh[j][i] = d[i] * g[i];
        dg[j][i] += d[i] * xh[j][i];
        db[j][i] += d[i];
        s1 += dxh[j][i];
        s2 += dxh[j][i] * xh[j][i];
      }
    }
    const float c1 = warp_sum(s1) / (float)p.h, c2 = warp_sum(s2) / (float)p.h;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (lane + 32 * j) * 8;
      float ds[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ds[i] = rstd * (dxh[j][i] - c1 - xh[j][i] * c2);
        dp[j][i] += ds[i];
        if (tt == 0) dt0[j][i] += ds[i]; else dt1[j][i] += ds[i];
      }
      if (id != 0 && p.dword) {  // padding_idx = 0
        float* w = p.dword + (long long)id * p.h + col;
        atomicAdd(reinterpret_cast<float4*>(w), make_float4(ds[0], ds[1], ds[2], ds[3]));
        atomicAdd(reinterpret_cast<float4*>(w + 4), make_float4(ds[4], ds[5], ds[6], ds[7]));
      }
    }
  }
  if (p.dgamma) flush_colsum<NV>(dg, ln_smem, p.dgamma, p.h, lane, warp);
  if (p.dbeta) flush_colsum<NV>(db, ln_smem, p.dbeta, p.h, lane, warp);
  if (p.dpos) flush_colsum<NV>(dp, ln_smem, p.dpos + (long long)pos * p.h, p.h, lane, warp);
  if (p.dtype) {
    flush_colsum<NV>(dt0, ln_smem, p.dtype, p.h, lane, warp);
    flush_colsum<NV>(dt1, ln_smem, p.dtype + p.h, p.h, lane, warp);
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// out[n] += sum_m x[m,n]   (bias gradients).  grid = (ceil(N/256), row chunks); a warp covers 256 columns.
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* x, long long ld, int m, int n, float* out,
                                                     int rows_per_block) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(m, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < n) {
    for (int r = r0 + warp; r < r1; r += 8) {
      float v[8];
      ld8(x + (long long)r * ld + col, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < n) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    atomicAdd(out + blockIdx.x * 256 + c, s);
  }
}

// ------------------------------------------------------------------------------------------------ casts and masks
struct CastSeg { const float* src; __nv_bfloat16* dst; long long n; };

__global__ void cast_multi_kernel(const CastSeg* segs, const int* block_seg, const long long* block_off) {
  const CastSeg s = segs[block_seg[blockIdx.x]];
  const long long base = block_off[blockIdx.x];
  const long long end = min(s.n, base + (long long)blockDim.x * 8 * 4);
  for (long long i = base + (long long)threadIdx.x * 8; i < end; i += (long long)blockDim.x * 8) {
    if (i + 8 <= s.n) {
      float v[8];
      ld8f(s.src + i, v);
      st8(s.dst + i, v);
    } else {
      for (long long k = i; k < s.n; ++k) s.dst[k] = __float2bfloat16_rn(s.src[k]);
    }
  }
}

// The fp32 source is read ONCE per step (ld.global.cs = evict-first): the refresh runs beside the first layers of the forward,
// whose activations live in L2.
__global__ void cast_kernel(const float* src, __nv_bfloat16* dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(src + i)), b = __ldcs(reinterpret_cast<const float4*>(src + i + 4));
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      st8(dst + i, v);
    } else {
      for (long long k = i; k < n; ++k) dst[k] = __float2bfloat16_rn(src[k]);
    }
  }
}

// bf16 -> fp32 (gradient buckets coming back from the bf16 all-reduce)
// Streaming accesses (ld.global.cs / st.global.cs = evict-first): 0.5 GB read once and 1 GB written once per step while the
// backward's GEMMs live on L2-resident activations.  Measured at 2 GPUs: no difference to plain accesses (5.49 ms/step both
// ways, VB_WIDEN_PLAIN=1 for the A/B run) -- the 0.33 ms this pass costs the step is HBM bandwidth taken from the GEMMs' weight
// streams, not L2 eviction; the hint stays because it is the honest description of the access pattern.
__global__ void uncast_kernel(const __nv_bfloat16* src, float* dst, long long n, int streaming) {
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n && streaming) {
      const uint4 u = __ldcs(reinterpret_cast<const uint4*>(src + i));
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      __stcs(reinterpret_cast<float4*>(dst + i), make_float4(a.x, a.y, b.x, b.y));
      __stcs(reinterpret_cast<float4*>(dst + i + 4), make_float4(c.x, c.y, d.x, d.y));
    } else if (i + 8 <= n) {
      float v[8];
      ld8(src + i, v);
      *reinterpret_cast<float4*>(dst + i) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dst + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      for (long long k = i; k < n; ++k) dst[k] = __bfloat162float(src[k]);
    }
  }
}

// additive attention mask, bit-exact restatement of (1.0 - m) * -10000.0 (vilbert_facebook_arch.py:530-540)
template <typename T>
__global__ void mask_bias_kernel(const T* m, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (1.0f - static_cast<float>(m[i])) * -10000.0f;
}

__global__ void i64_to_i32_kernel(const long long* src, int* dst, int n, int lo, int hi, int* err_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const long long v = src[i];
    if (v < lo || v >= hi) { if (err_flag) atomicExch(err_flag, 1); dst[i] = lo; }
    else dst[i] = (int)v;
  }
}

// ------------------------------------------------------------------------------------------------ element-wise
// y = dropout(x)  (used on both activations and their gradients: same site -> same mask)
__global__ void dropout_kernel(const __nv_bfloat16* x, __nv_bfloat16* y, long long n, float p, uint32_t site,
                               const unsigned long long* seed_ptr) {
  const uint64_t seed = *seed_ptr;
  const uint32_t thr = dropout_threshold(p);
  const float inv = 1.f / (1.f - p);
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    float v[8];
    bool k[8];
    ld8(x + i, v);
    keep8(seed, site, (uint32_t)i, thr, k);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = k[j] ? v[j] * inv : 0.f;
    st8(y + i, v);
  }
}

// dx = dy * act'(.) expressed through the activation OUTPUT y: tanh' = 1 - y^2, relu' = [y > 0]
__global__ void act_bwd_kernel(const __nv_bfloat16* dy, const __nv_bfloat16* y, __nv_bfloat16* dx, long long n, int act) {
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    float d[8], o[8];
    ld8(dy + i, d);
    ld8(y + i, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = (act == VB_ACT_TANH) ? d[j] * (1.f - o[j] * o[j]) : (o[j] > 0.f ? d[j] : 0.f);
    st8(dx + i, d);
  }
}

// ------------------------------------------------------------------------------------------------ visual location term
// loc_emb[m, n] = sum_j loc[m, j] * W[n, j] + b[n]   (image_location_embeddings, vilbert_facebook_arch.py:101-102; K = 5)
__global__ void loc_embed_fwd_kernel(const float* loc, const float* w, const float* b, __nv_bfloat16* out, int m, int n,
                                     int kdim) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const int row = blockIdx.y;
  if (col >= n || row >= m) return;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float acc = 0.f;
    for (int j = 0; j < kdim; ++j) acc += loc[row * kdim + j] * w[(col + i) * kdim + j];
    o[i] = acc + b[col + i];
  }
  st8(out + (long long)row * n + col, o);
}
// dW[n, j] += sum_m ds[m, n] * loc[m, j];  db[n] += sum_m ds[m, n]      (kdim <= 8)
__global__ void loc_embed_bwd_kernel(const __nv_bfloat16* ds, const float* loc, float* dw, float* db, int m, int n,
                                     int kdim, int rows_per_block) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(m, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float accb = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float g = __bfloat162float(ds[(long long)r * n + col]);
    accb += g;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < kdim) acc[j] += g * loc[r * kdim + j];
  }
  if (db) atomicAdd(db + col, accb);
  for (int j = 0; j < kdim; ++j) atomicAdd(dw + col * kdim + j, acc[j]);
}

// ------------------------------------------------------------------------------------------------ classifier tail + CE
// logits = h W^T + b  (Linear(1024, num_labels), vilbert_facebook_arch.py:577), CrossEntropyLoss mean (:637-639).
// One block; fp32 master weights are read directly (C <= 8 rows).
__global__ void __launch_bounds__(1024) cls_fwd_kernel(const __nv_bfloat16* h, const float* w, const float* bias,
                                                      const int* labels, float* logits, float* probs, float* loss,
                                                      int bsz, int kdim, int c) {
  __shared__ float s_logits[8192];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // one warp per (sample, class) dot product, 32 of them in flight (the 8-warp version walked the 32 outputs of the bs-16 /
  // 2-class head in four dependent rounds: 15 us at the END of the forward's critical path)
  const int nwarps = blockDim.x >> 5;
  for (int o = warp; o < bsz * c; o += nwarps) {
    const int bi = o / c, ci = o % c;
    float acc = 0.f;
#pragma unroll 4
    for (int k = lane; k < kdim; k += 32) acc += __bfloat162float(h[(long long)bi * kdim + k]) * w[(long long)ci * kdim + k];
    acc = warp_sum(acc);
    if (lane == 0) { s_logits[o] = acc + bias[ci]; logits[o] = acc + bias[ci]; }
  }
  __syncthreads();
  if (warp == 0) {
    float l = 0.f, nv = 0.f;
    for (int bi = lane; bi < bsz; bi += 32) {
      float mx = -INFINITY;
      for (int ci = 0; ci < c; ++ci) mx = fmaxf(mx, s_logits[bi * c + ci]);
      float se = 0.f;
      for (int ci = 0; ci < c; ++ci) se += expf(s_logits[bi * c + ci] - mx);
      const float lse = mx + logf(se);
      for (int ci = 0; ci < c; ++ci) probs[bi * c + ci] = expf(s_logits[bi * c + ci] - lse);
      // nn.CrossEntropyLoss() defaults: ignore_index = -100, mean over the samples that are NOT ignored
      if (labels && labels[bi] != VB_IGNORE_INDEX) { l += lse - s_logits[bi * c + labels[bi]]; nv += 1.f; }
    }
    l = warp_sum(l);
    nv = warp_sum(nv);
    if (lane == 0 && loss) *loss = labels ? l / nv : 0.f;     // every label ignored -> 0/0 = NaN, as torch
  }
}

// dlogits = dloss * (probs - onehot) / B + dlogits_ext;  dW = dlogits^T h;  db = colsum(dlogits);  dh = dlogits W
__global__ void __launch_bounds__(256) cls_bwd_kernel(const __nv_bfloat16* h, const float* w, const int* labels,
                                                      const float* probs, const float* dloss, const float* dlogits_ext,
                                                      float* dw, float* db, __nv_bfloat16* dh, int bsz, int kdim, int c) {
  __shared__ float s_dl[8192];
  __shared__ int s_nv;
  const float gl = (dloss && labels) ? *dloss : 0.f;
  if (threadIdx.x == 0) s_nv = 0;
  __syncthreads();
  if (labels) {
    int mine = 0;
    for (int bi = threadIdx.x; bi < bsz; bi += blockDim.x) mine += labels[bi] != VB_IGNORE_INDEX ? 1 : 0;
    if (mine) atomicAdd(&s_nv, mine);
  }
  __syncthreads();
  const float inv_nv = 1.f / (float)s_nv;
  for (int o = threadIdx.x; o < bsz * c; o += blockDim.x) {
    const int bi = o / c, ci = o % c;
    float g = 0.f;
    if (labels && labels[bi] != VB_IGNORE_INDEX) g = gl * (probs[o] - (labels[bi] == ci ? 1.f : 0.f)) * inv_nv;
    if (dlogits_ext) g += dlogits_ext[o];
    s_dl[o] = g;
  }
  __syncthreads();
  // the hidden dimension is spread over the GRID (every block recomputes the tiny dlogits above): one block of 256 threads
  // walked 4 x (c + 1) x bsz dependent loads = 30 us at the START of the backward's critical path
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < kdim; k += gridDim.x * blockDim.x) {
    for (int ci = 0; ci < c; ++ci) {
      float acc = 0.f;
#pragma unroll 8
      for (int bi = 0; bi < bsz; ++bi) acc += s_dl[bi * c + ci] * __bfloat162float(h[(long long)bi * kdim + k]);
      if (dw) dw[(long long)ci * kdim + k] = acc;
    }
#pragma unroll 4
    for (int bi = 0; bi < bsz; ++bi) {
      float acc = 0.f;
      for (int ci = 0; ci < c; ++ci) acc += s_dl[bi * c + ci] * w[(long long)ci * kdim + k];
      dh[(long long)bi * kdim + k] = __float2bfloat16_rn(acc);
    }
  }
  if (db && blockIdx.x == 0 && threadIdx.x < c) {
    float acc = 0.f;
    for (int bi = 0; bi < bsz; ++bi) acc += s_dl[bi * c + threadIdx.x];
    db[threadIdx.x] = acc;
  }
}


// ------------------------------------------------------------------------------------------------ batch staging
// ONE launch moves a whole batch from the caller's tensors into the plan's static buffers (the graphs read those addresses):
// ids / token types / labels (int64 or int32 -> int32, range-checked like nn.Embedding / CrossEntropyLoss check them), the
// two attention masks (-> additive bias, bit-exact (1 - m) * -10000), region features (fp32 -> bf16, or a bf16 copy) and
// the 5-d boxes (fp32 copy).  Blocks are dealt to the segments in order; a segment with n = 0 takes none.
struct StageSeg { const void* src; void* dst; long long n; int kind; int dtype; int lo, hi; int err_bit; int first_block; };
struct StageParams { StageSeg seg[VB_STAGE_MAX_SEGS]; int nseg; int* err_flag; };
constexpr int STAGE_PER_BLOCK = 256 * 8;

__global__ void __launch_bounds__(256) stage_batch_kernel(const StageParams p) {
  int si = 0;
#pragma unroll 1
  for (int i = 1; i < p.nseg; ++i) if ((int)blockIdx.x >= p.seg[i].first_block) si = i;
  const StageSeg& g = p.seg[si];
  const long long base = (long long)((int)blockIdx.x - g.first_block) * STAGE_PER_BLOCK;
  if (g.kind == VB_STAGE_FEAT && base + STAGE_PER_BLOCK <= g.n && (g.n & 7) == 0) {
    // vector path: 8 elements per thread
    const long long i = base + threadIdx.x * 8;
    if (g.dtype == VB_DT_F32) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(g.src) + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(g.src) + i) + 1);
      uint4 o;
      o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w); o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(g.dst) + i) = o;
    } else {
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(g.dst) + i) =
          __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(g.src) + i));
    }
    return;
  }
  bool bad = false;
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const long long i = base + j * 256 + threadIdx.x;
    if (i >= g.n) break;
    if (g.kind == VB_STAGE_INDEX) {
      const long long v = g.dtype == VB_DT_I64 ? static_cast<const long long*>(g.src)[i] : (long long)static_cast<const int*>(g.src)[i];
      const bool ignored = g.err_bit == VB_STAGE_ERR_LABEL && v == VB_IGNORE_INDEX;   // CrossEntropyLoss(ignore_index=-100)
      const bool ok = (v >= g.lo && v < g.hi) || ignored;
      bad |= !ok;
      static_cast<int*>(g.dst)[i] = ok ? (int)v : g.lo;     // clamped: nothing downstream may index out of bounds
    } else if (g.kind == VB_STAGE_MASK) {
      float m;
      if (g.dtype == VB_DT_I64) m = static_cast<float>(static_cast<const long long*>(g.src)[i]);
      else if (g.dtype == VB_DT_I32) m = static_cast<float>(static_cast<const int*>(g.src)[i]);
      else m = static_cast<const float*>(g.src)[i];
      static_cast<float*>(g.dst)[i] = (1.0f - m) * -10000.0f;
    } else if (g.kind == VB_STAGE_FEAT) {
      static_cast<__nv_bfloat16*>(g.dst)[i] = g.dtype == VB_DT_F32 ? __float2bfloat16_rn(static_cast<const float*>(g.src)[i])
                                                                   : static_cast<const __nv_bfloat16*>(g.src)[i];
    } else {
      static_cast<float*>(g.dst)[i] = static_cast<const float*>(g.src)[i];
    }
  }
  if (bad && p.err_flag) atomicOr(p.err_flag, g.err_bit);
}

__global__ void seed_advance_to_kernel(unsigned long long* seed, unsigned long long* snapshot) {
  const unsigned long long s = *seed * 6364136223846793005ull + 1442695040888963407ull;
  *seed = s;
  *snapshot = s;
}

}  // namespace vb

// =================================================================================================== C ABI
using namespace vb;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int fill_ln(LnParams& p, const vb_layernorm_args* a) {
  VB_REQUIRE(a != nullptr, "null args");
  VB_REQUIRE(a->x && a->gamma && a->mean && a->rstd, "x, gamma, mean, rstd are required");
  VB_REQUIRE(a->m > 0 && a->h > 0 && a->h % 256 == 0, "h must be a multiple of 256");
  VB_REQUIRE(a->ldx % 8 == 0 && a->ldres % 8 == 0 && a->ldy % 8 == 0 && a->lddy % 8 == 0 && a->lddx % 8 == 0 && a->lddres % 8 == 0,
             "leading dimensions must be multiples of 8 elements");
  VB_REQUIRE(aligned16(a->x) && aligned16(a->res) && aligned16(a->y) && aligned16(a->dy) && aligned16(a->dx) && aligned16(a->dres) &&
             aligned16(a->gamma) && aligned16(a->beta), "pointers must be 16-byte aligned");
  VB_REQUIRE(a->p_in >= 0.f && a->p_in < 1.f && a->p_out >= 0.f && a->p_out < 1.f, "dropout p in [0,1)");
  VB_REQUIRE((a->p_in == 0.f && a->p_out == 0.f) || a->seed != nullptr, "dropout needs a device seed pointer");
  p.x = (const __nv_bfloat16*)a->x; p.res = (const __nv_bfloat16*)a->res; p.gamma = a->gamma; p.beta = a->beta;
  p.y = (__nv_bfloat16*)a->y; p.mean = a->mean; p.rstd = a->rstd;
  p.res32 = a->res_f32; p.y32 = a->y_f32; p.ldres32 = a->ldres_f32; p.ldy32 = a->ldy_f32;
  VB_REQUIRE(aligned16(a->res_f32) && aligned16(a->y_f32) && a->ldres_f32 % 4 == 0 && a->ldy_f32 % 4 == 0, "fp32 residual / output alignment");
  p.dy = (const __nv_bfloat16*)a->dy; p.dx = (__nv_bfloat16*)a->dx; p.dres = (__nv_bfloat16*)a->dres;
  p.dgamma = a->dgamma; p.dbeta = a->dbeta; p.dbias = a->dbias;
  p.ldx = a->ldx; p.ldres = a->ldres; p.ldy = a->ldy; p.lddy = a->lddy; p.lddx = a->lddx; p.lddres = a->lddres;
  p.m = a->m; p.h = a->h; p.eps = a->eps; p.p_in = a->p_in; p.p_out = a->p_out;
  p.site_in = a->site_in; p.site_out = a->site_out; p.seed = (const unsigned long long*)a->seed;
  return VB_OK;
}

extern "C" int vb_layernorm_fwd(const vb_layernorm_args* a, void* stream) {
  LnParams p;
  int rc = fill_ln(p, a);
  if (rc != VB_OK) return rc;
  VB_REQUIRE(a->y && a->beta, "y and beta are required");
  cudaStream_t s = (cudaStream_t)stream;
  rc = dispatch_nv(p.h, [&](auto nv) {
    ln_fwd_kernel<decltype(nv)::value><<<ln_grid(p.m, getenv("VB_LN_FWD_ROWS") ? atoi(getenv("VB_LN_FWD_ROWS")) : 1), LN_WARPS * 32, 0, s>>>(p);
    return VB_OK;
  });
  if (rc != VB_OK) return rc;
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_layernorm_bwd(const vb_layernorm_args* a, void* stream) {
  LnParams p;
  int rc = fill_ln(p, a);
  if (rc != VB_OK) return rc;
  VB_REQUIRE(a->dy && (a->dx || a->dres), "dy and at least one of dx / dres are required");
  cudaStream_t s = (cudaStream_t)stream;
  rc = dispatch_nv(p.h, [&](auto nv) {
    auto kern = ln_bwd_kernel<decltype(nv)::value>;
    const int smem = LN_WARPS * p.h * (int)sizeof(float);
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_WARPS * 2048 * (int)sizeof(float));
      attr_set = true;
    }
    // one row per warp up to two full waves of blocks, then rows are strided over the grid
    // four rows per warp: a quarter of the blocks issue a quarter of the column-sum atomics, and the smaller grid leaves SMs to
    // the kernels of the other streams (alone the kernel is slower, 9.4 -> 11.3 us; the step is faster, 5.71 -> 5.51 ms)
    int grid = ln_grid(p.m, getenv("VB_LN_BWD_ROWS") ? atoi(getenv("VB_LN_BWD_ROWS")) : 4);
    if (grid > 2 * 148) grid = ln_grid(p.m, (p.m + 2 * 148 * LN_WARPS - 1) / (2 * 148 * LN_WARPS));
    kern<<<grid, LN_WARPS * 32, smem, s>>>(p);
    return VB_OK;
  });
  if (rc != VB_OK) return rc;
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

static int fill_emb(EmbParams& p, const vb_embed_args* a) {
  VB_REQUIRE(a != nullptr, "null args");
  VB_REQUIRE(a->ids && a->word && a->pos && a->type && a->gamma && a->mean && a->rstd, "missing pointer");
  VB_REQUIRE(a->b > 0 && a->t > 0 && a->h % 256 == 0, "h must be a multiple of 256");
  VB_REQUIRE(a->p_out == 0.f || a->seed != nullptr, "dropout needs a device seed pointer");
  p.ids = a->ids; p.type_ids = a->type_ids; p.word = a->word; p.pos = a->pos; p.type = a->type;
  p.gamma = a->gamma; p.beta = a->beta; p.y = (__nv_bfloat16*)a->y; p.y32 = a->y_f32; p.mean = a->mean; p.rstd = a->rstd;
  p.dy = (const __nv_bfloat16*)a->dy; p.dword = a->dword; p.dpos = a->dpos; p.dtype = a->dtype;
  p.dgamma = a->dgamma; p.dbeta = a->dbeta;
  p.b = a->b; p.t = a->t; p.h = a->h; p.vocab = a->vocab; p.eps = a->eps; p.p_out = a->p_out; p.site_out = a->site_out;
  p.seed = (const unsigned long long*)a->seed;
  return VB_OK;
}

extern "C" int vb_embed_text_fwd(const vb_embed_args* a, void* stream) {
  EmbParams p;
  int rc = fill_emb(p, a);
  if (rc != VB_OK) return rc;
  VB_REQUIRE(a->y && a->beta, "y and beta are required");
  cudaStream_t s = (cudaStream_t)stream;
  rc = dispatch_nv(p.h, [&](auto nv) {
    emb_fwd_kernel<decltype(nv)::value><<<ln_grid(p.b * p.t, 1), LN_WARPS * 32, 0, s>>>(p);
    return VB_OK;
  });
  if (rc != VB_OK) return rc;
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_embed_text_bwd(const vb_embed_args* a, void* stream) {
  EmbParams p;
  int rc = fill_emb(p, a);
  if (rc != VB_OK) return rc;
  VB_REQUIRE(a->dy, "dy is required");
  cudaStream_t s = (cudaStream_t)stream;
  rc = dispatch_nv(p.h, [&](auto nv) {
    emb_bwd_kernel<decltype(nv)::value><<<p.t, LN_WARPS * 32, LN_WARPS * p.h * sizeof(float), s>>>(p);
    return VB_OK;
  });
  if (rc != VB_OK) return rc;
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_colsum_bf16(const void* x, int64_t ld, int32_t m, int32_t n, float* out, void* stream) {
  VB_REQUIRE(x && out && m > 0 && n > 0 && n % 8 == 0 && ld % 8 == 0 && aligned16(x), "bad colsum arguments");
  const int rows_per_block = 128;
  dim3 grid((n + 255) / 256, (m + rows_per_block - 1) / rows_per_block);
  colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, m, n, out, rows_per_block);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  VB_REQUIRE(src && dst && n > 0 && aligned16(src) && aligned16(dst), "bad cast arguments");
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  static const int cap = getenv("VB_CAST_BLOCKS") ? atoi(getenv("VB_CAST_BLOCKS")) : 148 * 16;
  if (blocks > cap) blocks = cap;
  cast_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream) {
  VB_REQUIRE(src && dst && n > 0 && aligned16(src) && aligned16(dst), "bad cast arguments");
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  static const int streaming = getenv("VB_WIDEN_PLAIN") && atoi(getenv("VB_WIDEN_PLAIN")) ? 0 : 1;
  uncast_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, dst, n, streaming);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_cast_f32_bf16_multi(const void* segs, const int32_t* block_seg, const int64_t* block_off,
                                      int32_t num_blocks, void* stream) {
  VB_REQUIRE(segs && block_seg && block_off && num_blocks > 0, "bad multi-cast arguments");
  cast_multi_kernel<<<num_blocks, 256, 0, (cudaStream_t)stream>>>((const CastSeg*)segs, block_seg,
                                                                 (const long long*)block_off);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_mask_bias(const void* mask, int32_t mask_dtype, float* out, int32_t n, void* stream) {
  VB_REQUIRE(mask && out && n > 0, "bad mask arguments");
  const int blocks = (n + 255) / 256;
  cudaStream_t s = (cudaStream_t)stream;
  switch (mask_dtype) {
    case VB_DT_I64: mask_bias_kernel<long long><<<blocks, 256, 0, s>>>((const long long*)mask, out, n); break;
    case VB_DT_I32: mask_bias_kernel<int><<<blocks, 256, 0, s>>>((const int*)mask, out, n); break;
    case VB_DT_F32: mask_bias_kernel<float><<<blocks, 256, 0, s>>>((const float*)mask, out, n); break;
    default: vb_set_last_error("mask dtype", "supported: i64, i32, f32"); return VB_ERR_UNSUPPORTED;
  }
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_i64_to_i32(const int64_t* src, int32_t* dst, int32_t n, int32_t lo, int32_t hi, int32_t* err_flag,
                             void* stream) {
  VB_REQUIRE(src && dst && n > 0, "bad arguments");
  i64_to_i32_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const long long*)src, dst, n, lo, hi, err_flag);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_dropout_bf16(const void* x, void* y, int64_t n, float p, uint32_t site, const uint64_t* seed,
                               void* stream) {
  VB_REQUIRE(x && y && seed && n > 0 && n % 8 == 0 && p > 0.f && p < 1.f && aligned16(x) && aligned16(y), "bad dropout arguments");
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  dropout_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, p, site,
                                                               (const unsigned long long*)seed);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_act_bwd_bf16(const void* dy, const void* y, void* dx, int64_t n, int32_t act, void* stream) {
  VB_REQUIRE(dy && y && dx && n > 0 && n % 8 == 0 && (act == VB_ACT_TANH || act == VB_ACT_RELU), "bad act_bwd arguments");
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  act_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y,
                                                               (__nv_bfloat16*)dx, n, act);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_loc_embed_fwd(const float* loc, const float* w, const float* b, void* out, int32_t m, int32_t n,
                                int32_t kdim, void* stream) {
  VB_REQUIRE(loc && w && b && out && m > 0 && n % 8 == 0 && kdim > 0 && kdim <= 8, "bad loc_embed arguments");
  dim3 grid((n / 8 + 127) / 128, m);
  loc_embed_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(loc, w, b, (__nv_bfloat16*)out, m, n, kdim);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_loc_embed_bwd(const void* ds, const float* loc, float* dw, float* db, int32_t m, int32_t n,
                                int32_t kdim, void* stream) {
  VB_REQUIRE(ds && loc && dw && m > 0 && n > 0 && kdim > 0 && kdim <= 8, "bad loc_embed arguments");
  const int rows_per_block = 64;
  dim3 grid((n + 127) / 128, (m + rows_per_block - 1) / rows_per_block);
  loc_embed_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)ds, loc, dw, db, m, n, kdim,
                                                              rows_per_block);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_cls_ce_fwd(const void* h, const float* w, const float* bias, const int32_t* labels, float* logits,
                             float* probs, float* loss, int32_t bsz, int32_t kdim, int32_t c, void* stream) {
  VB_REQUIRE(h && w && bias && logits && probs && bsz > 0 && bsz * c <= 8192 && c <= 8 && kdim > 0, "bad classifier arguments (B*C <= 8192, C <= 8)");
  cls_fwd_kernel<<<1, (bsz * c >= 32) ? 1024 : 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)h, w, bias, labels, logits, probs, loss, bsz, kdim, c);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_cls_ce_bwd(const void* h, const float* w, const int32_t* labels, const float* probs,
                             const float* dloss, const float* dlogits_ext, float* dw, float* db, void* dh, int32_t bsz,
                             int32_t kdim, int32_t c, void* stream) {
  VB_REQUIRE(h && w && probs && dh && bsz > 0 && bsz * c <= 8192 && c <= 8 && kdim > 0, "bad classifier arguments (B*C <= 8192, C <= 8)");
  cls_bwd_kernel<<<(kdim + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)h, w, labels, probs, dloss, dlogits_ext, dw, db,
                                                      (__nv_bfloat16*)dh, bsz, kdim, c);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

// dropout seed stream: advanced once per training forward INSIDE the captured graph, so every replay draws new masks
__global__ void seed_advance_kernel(unsigned long long* seed) {
  *seed = *seed * 6364136223846793005ull + 1442695040888963407ull;
}
extern "C" int vb_stage_batch(const vb_stage_seg* segs, int32_t nseg, int32_t* err_flag, void* stream) {
  VB_REQUIRE(segs != nullptr && nseg > 0 && nseg <= VB_STAGE_MAX_SEGS, "1..VB_STAGE_MAX_SEGS segments");
  StageParams p;
  p.nseg = nseg;
  p.err_flag = err_flag;
  int blocks = 0;
  for (int i = 0; i < nseg; ++i) {
    const vb_stage_seg& g = segs[i];
    VB_REQUIRE(g.n >= 0 && (g.n == 0 || (g.src && g.dst)), "segment pointers");
    VB_REQUIRE(g.kind >= VB_STAGE_INDEX && g.kind <= VB_STAGE_COPY_F32, "segment kind");
    VB_REQUIRE(g.kind != VB_STAGE_INDEX || g.dtype == VB_DT_I64 || g.dtype == VB_DT_I32, "index segments are int64 or int32");
    VB_REQUIRE(g.kind != VB_STAGE_FEAT || g.dtype == VB_DT_F32 || g.dtype == VB_DT_BF16, "feature segments are fp32 or bf16");
    VB_REQUIRE(g.kind != VB_STAGE_FEAT || (aligned16(g.src) && aligned16(g.dst)), "feature segments must be 16-byte aligned");
    p.seg[i].src = g.src; p.seg[i].dst = g.dst; p.seg[i].n = g.n; p.seg[i].kind = g.kind; p.seg[i].dtype = g.dtype;
    p.seg[i].lo = g.lo; p.seg[i].hi = g.hi; p.seg[i].err_bit = g.err_bit; p.seg[i].first_block = blocks;
    blocks += static_cast<int>((g.n + STAGE_PER_BLOCK - 1) / STAGE_PER_BLOCK);
  }
  VB_REQUIRE(blocks > 0, "nothing to stage");
  stage_batch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_seed_advance_to(uint64_t* seed, uint64_t* snapshot, void* stream) {
  VB_REQUIRE(seed != nullptr && snapshot != nullptr, "null seed");
  seed_advance_to_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)seed, (unsigned long long*)snapshot);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_seed_advance(uint64_t* seed, void* stream) {
  VB_REQUIRE(seed != nullptr, "null seed");
  seed_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)seed);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}
