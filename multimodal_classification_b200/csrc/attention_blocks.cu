// Joining kernels for attention over sequences above 128 (BASELINE config 4: 257 DINOv2 patch tokens as regions).  The fused
// tcgen05 kernel of attention.cu handles one block of <= 128 queries x <= 128 keys per CTA; the host loops over blocks
// (ops.py) and these HBM-bound kernels apply the flash-attention algebra between them:
//   forward   O = sum_j exp(lse_j - LSE) O_j,  LSE = log sum_j exp(lse_j)                       (attn_merge_kernel)
//   backward  Delta = rowsum(dO * O) = sum over ALL keys of P dP, handed to every key block     (attn_delta_kernel)
//             dQ = sum over key blocks, dK / dV = sum over query blocks of the blocks' partials (sum_rows_kernel)
// One warp per (sample, head, query row); 16-byte accesses along the head width.
#include "common.cuh"
#include "../../include/vilbert_b200.h"

namespace vb {

constexpr int MAX_PARTS = 4;
struct PartPtrs { const __nv_bfloat16* o[MAX_PARTS]; const float* lse[MAX_PARTS]; };

// D = 64: lanes 0..7 hold 8 columns each; D = 128: lanes 0..15.
__global__ void __launch_bounds__(256)
attn_merge_kernel(const PartPtrs parts, int n_parts, long long ldp, __nv_bfloat16* __restrict__ out, long long ldo,
                  float* __restrict__ lse_out, int batch, int heads, int sq, long long batch_rows, int d) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)batch * heads * sq;
  if (w >= total) return;
  const int row = (int)(w % sq);
  const int h = (int)((w / sq) % heads);
  const int b = (int)(w / ((long long)sq * heads));
  const long long li = ((long long)b * heads + h) * 128 + row;
  float l[MAX_PARTS], m = -INFINITY;
#pragma unroll
  for (int j = 0; j < MAX_PARTS; ++j) {
    l[j] = j < n_parts ? parts.lse[j][li] : -INFINITY;
    m = fmaxf(m, l[j]);
  }
  float wgt[MAX_PARTS], tot = 0.f;
#pragma unroll
  for (int j = 0; j < MAX_PARTS; ++j) { wgt[j] = j < n_parts ? expf(l[j] - m) : 0.f; tot += wgt[j]; }
  if (lane == 0) lse_out[li] = m + logf(tot);
  if (lane * 8 >= d) return;
  const float inv = 1.f / tot;
  const long long off = ((long long)b * batch_rows + row);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < MAX_PARTS; ++j) {
    if (j < n_parts) {
      const uint4 u = *reinterpret_cast<const uint4*>(parts.o[j] + off * ldp + h * d + lane * 8);
      const float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
      const float s = wgt[j] * inv;
      acc[0] += s * a.x; acc[1] += s * a.y; acc[2] += s * bb.x; acc[3] += s * bb.y;
      acc[4] += s * c.x; acc[5] += s * c.y; acc[6] += s * e.x; acc[7] += s * e.y;
    }
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]); o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(out + off * ldo + h * d + lane * 8) = o;
}

__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ out, long long ldo, const __nv_bfloat16* __restrict__ dout, long long lddo,
                  float* __restrict__ delta, int batch, int heads, int sq, long long batch_rows, int d) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)batch * heads * sq;
  if (w >= total) return;
  const int row = (int)(w % sq);
  const int h = (int)((w / sq) % heads);
  const int b = (int)(w / ((long long)sq * heads));
  const long long off = ((long long)b * batch_rows + row);
  float part = 0.f;
  if (lane * 8 < d) {
    const uint4 u = *reinterpret_cast<const uint4*>(out + off * ldo + h * d + lane * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(dout + off * lddo + h * d + lane * 8);
    const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
    const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y), g2 = unpack_bf16x2(g.z), g3 = unpack_bf16x2(g.w);
    part = a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y + a3.x * g3.x + a3.y * g3.y;
  }
  part = warp_sum(part);
  if (lane == 0) delta[((long long)b * heads + h) * 128 + row] = part;
}

struct SumPtrs { const __nv_bfloat16* p[MAX_PARTS]; };

__global__ void __launch_bounds__(256)
sum_rows_kernel(const SumPtrs parts, int n_parts, long long ldp, __nv_bfloat16* __restrict__ dst, long long ldd, long long rows,
                int width) {
  const int wv = width >> 3;
  const long long total = rows * wv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / wv;
    const int c = (int)(idx % wv) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < MAX_PARTS; ++j) {
      if (j < n_parts) {
        const uint4 u = *reinterpret_cast<const uint4*>(parts.p[j] + r * ldp + c);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), e = unpack_bf16x2(u.z), f = unpack_bf16x2(u.w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += e.x; acc[5] += e.y; acc[6] += f.x; acc[7] += f.y;
      }
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]); o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(dst + r * ldd + c) = o;
  }
}

static bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace vb

extern "C" int vb_attn_merge(const void* const* o_parts, const float* const* lse_parts, int32_t n_parts, int64_t ldp, void* out,
                             int64_t ldo, float* lse_out, int32_t batch, int32_t heads, int32_t sq, int64_t batch_rows, int32_t d,
                             void* stream) {
  VB_REQUIRE(o_parts && lse_parts && out && lse_out && n_parts >= 1 && n_parts <= vb::MAX_PARTS, "1..4 parts");
  VB_REQUIRE(batch > 0 && heads > 0 && sq >= 1 && sq <= 128 && batch_rows >= sq && (d == 64 || d == 128), "bad geometry");
  VB_REQUIRE(ldp % 8 == 0 && ldo % 8 == 0 && vb::al16p(out), "row strides must be multiples of 8, pointers 16-byte aligned");
  vb::PartPtrs pp;
  for (int j = 0; j < vb::MAX_PARTS; ++j) {
    pp.o[j] = j < n_parts ? static_cast<const __nv_bfloat16*>(o_parts[j]) : nullptr;
    pp.lse[j] = j < n_parts ? lse_parts[j] : nullptr;
    VB_REQUIRE(j >= n_parts || (o_parts[j] && lse_parts[j] && vb::al16p(o_parts[j])), "null or misaligned part");
  }
  const long long warps = (long long)batch * heads * sq;
  vb::attn_merge_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pp, n_parts, ldp, static_cast<__nv_bfloat16*>(out), ldo, lse_out, batch, heads, sq, batch_rows, d);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_attn_delta(const void* out, int64_t ldo, const void* dout, int64_t lddo, float* delta, int32_t batch,
                             int32_t heads, int32_t sq, int64_t batch_rows, int32_t d, void* stream) {
  VB_REQUIRE(out && dout && delta && batch > 0 && heads > 0 && sq >= 1 && sq <= 128 && batch_rows >= sq && (d == 64 || d == 128),
             "bad arguments");
  VB_REQUIRE(ldo % 8 == 0 && lddo % 8 == 0 && vb::al16p(out) && vb::al16p(dout), "row strides must be multiples of 8, pointers 16-byte aligned");
  const long long warps = (long long)batch * heads * sq;
  vb::attn_delta_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(out), ldo, static_cast<const __nv_bfloat16*>(dout), lddo, delta, batch, heads, sq,
      batch_rows, d);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_sum_rows_bf16(const void* const* parts, int32_t n_parts, int64_t ldp, void* dst, int64_t ldd, int64_t rows,
                                int32_t width, void* stream) {
  VB_REQUIRE(parts && dst && n_parts >= 1 && n_parts <= vb::MAX_PARTS && rows > 0 && width > 0 && width % 8 == 0, "bad arguments");
  VB_REQUIRE(ldp % 8 == 0 && ldd % 8 == 0 && vb::al16p(dst), "row strides must be multiples of 8, pointers 16-byte aligned");
  vb::SumPtrs sp;
  for (int j = 0; j < vb::MAX_PARTS; ++j) {
    sp.p[j] = j < n_parts ? static_cast<const __nv_bfloat16*>(parts[j]) : nullptr;
    VB_REQUIRE(j >= n_parts || (parts[j] && vb::al16p(parts[j])), "null or misaligned part");
  }
  long long blocks = (rows * (width / 8) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  vb::sum_rows_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(sp, n_parts, ldp,
                                                                                      static_cast<__nv_bfloat16*>(dst), ldd, rows, width);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}
