// DINOv2 multi-layer fusion tail (reference models/feature_extractors/dinov2_multilayer.py:342-381), the part of that
// extractor that is not the third-party ViT: concatenate the patch tokens of L transformer layers along the feature axis,
// view them as a g x g grid, resize bilinearly to t x t regions (F.interpolate, align_corners=False) and project
// (Linear -> LayerNorm -> GELU -> Linear; the two Linears run on vb_gemm_bf16, the LayerNorm on vb_layernorm_fwd).
// Here: the fused concat + resize gather and the erf-GELU between LayerNorm and the second Linear.
#include "common.cuh"
#include "../../include/vilbert_b200.h"

namespace vb {

struct LayerPtrs { const float* p[8]; };

// out[(b*t*t + oy*t + ox), l*h + c] = bilinear(layer_l[b, tok0 + y*g + x, c]) ; one thread = 8 consecutive channels
__global__ void bilinear_concat_kernel(const LayerPtrs layers, int num_layers, __nv_bfloat16* __restrict__ out, int batch, int g,
                                       int t, int h, long long batch_stride, long long token_stride, int tok0) {
  const int hv = h >> 3;
  const long long total = (long long)batch * t * t * num_layers * hv;
  const float scale = (float)g / (float)t;      // torch: area_pixel_compute_scale (align_corners=False, no scale_factor)
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % hv);
    long long r = idx / hv;
    const int l = (int)(r % num_layers); r /= num_layers;
    const int ox = (int)(r % t); r /= t;
    const int oy = (int)(r % t);
    const int b = (int)(r / t);
    // torch area_pixel_compute_source_index: max(scale * (dst + 0.5) - 0.5, 0)
    const float sy = fmaxf(scale * ((float)oy + 0.5f) - 0.5f, 0.f), sx = fmaxf(scale * ((float)ox + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < g - 1 ? 1 : 0), x1 = x0 + (x0 < g - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* base = layers.p[l] + (long long)b * batch_stride + (long long)tok0 * token_stride + cv * 8;
    float o[8];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(base + (long long)(y0 * g + x0) * token_stride + half * 4));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(base + (long long)(y0 * g + x1) * token_stride + half * 4));
      const float4 c = __ldg(reinterpret_cast<const float4*>(base + (long long)(y1 * g + x0) * token_stride + half * 4));
      const float4 d = __ldg(reinterpret_cast<const float4*>(base + (long long)(y1 * g + x1) * token_stride + half * 4));
      o[half * 4 + 0] = hy * (hx * a.x + lx * bb.x) + ly * (hx * c.x + lx * d.x);
      o[half * 4 + 1] = hy * (hx * a.y + lx * bb.y) + ly * (hx * c.y + lx * d.y);
      o[half * 4 + 2] = hy * (hx * a.z + lx * bb.z) + ly * (hx * c.z + lx * d.z);
      o[half * 4 + 3] = hy * (hx * a.w + lx * bb.w) + ly * (hx * c.w + lx * d.w);
    }
    uint4 u;
    u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]); u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
    const long long row = ((long long)b * t + oy) * t + ox;
    *reinterpret_cast<uint4*>(out + row * ((long long)num_layers * h) + (long long)l * h + cv * 8) = u;
  }
}

__global__ void gelu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + i);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    uint4 o;
    o.x = pack_bf16x2(gelu_erf(a.x), gelu_erf(a.y)); o.y = pack_bf16x2(gelu_erf(b.x), gelu_erf(b.y));
    o.z = pack_bf16x2(gelu_erf(c.x), gelu_erf(c.y)); o.w = pack_bf16x2(gelu_erf(d.x), gelu_erf(d.y));
    *reinterpret_cast<uint4*>(y + i) = o;
  }
}

}  // namespace vb

extern "C" int vb_bilinear_concat(const float* const* layers, int32_t num_layers, void* out, int32_t batch, int32_t grid, int32_t target,
                                  int32_t h, int64_t batch_stride, int64_t token_stride, int32_t first_token, void* stream) {
  VB_REQUIRE(layers && out && num_layers >= 1 && num_layers <= 8, "1..8 layers");
  VB_REQUIRE(batch > 0 && grid > 0 && target > 0 && h > 0 && h % 8 == 0 && token_stride % 4 == 0 && batch_stride % 4 == 0, "bad geometry");
  vb::LayerPtrs lp;
  for (int i = 0; i < 8; ++i) {
    lp.p[i] = i < num_layers ? layers[i] : nullptr;
    VB_REQUIRE(i >= num_layers || (layers[i] != nullptr && (reinterpret_cast<uintptr_t>(layers[i]) & 15) == 0), "layer pointers must be 16-byte aligned");
  }
  const long long total = (long long)batch * target * target * num_layers * (h / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  vb::bilinear_concat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(lp, num_layers, (__nv_bfloat16*)out, batch, grid, target, h,
                                                                             batch_stride, token_stride, first_token);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_gelu_bf16(const void* x, void* y, int64_t n, void* stream) {
  VB_REQUIRE(x && y && n > 0 && n % 8 == 0, "n must be a positive multiple of 8");
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  vb::gelu_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}
