// Bandwidth kernels of the ResNet-152 RoI feature stage (reference models/feature_extractors/resnet152_roi.py:49-74,
// 145-178).  Activations are NHWC bf16 so that every convolution is a tcgen05 GEMM over [pixels, channels]
// (vb_gemm_bf16 with the folded-BatchNorm scale/bias, residual aux and ReLU epilogue); the kernels here move data
// between those GEMMs with 16-byte coalesced accesses (8 channels per thread):
//   stem im2col (fp32 NCHW image -> bf16 [pixels, 7*7*3 padded to 152])        conv1 7x7/2          :49 (resnet.conv1)
//   im2col (NHWC -> [pixels, kh*kw*C]) for the 3x3 and the strided 1x1 convolutions of the bottlenecks
//   3x3/2 max-pool                                                              resnet.maxpool
//   RoIPool (quantised max pooling, torchvision.ops.RoIPool semantics)          :126, :167-170   -- bit-exact index maths
//   RoIAlign (bilinear, sampling_ratio, torchvision.ops.roi_align semantics)    fasterrcnn_resnet152.py:130-134
//   global average pool over the 7x7 (or any) spatial extent -> fp32            :71-73
#include <cfloat>

#include "common.cuh"
#include "../../include/vilbert_b200.h"

namespace vb {

__device__ __forceinline__ void ld8b(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

// ------------------------------------------------------------------------------------------------ stem im2col
// y[(n*ho + oy)*wo + ox, (ky*kw + kx)*3 + c] = img[n, c, oy*stride - pad + ky, ox*stride - pad + kx]  (0 outside), columns
// [kh*kw*3, kpad) are zero.  One block per (image, output row, segment of STEM_TW output pixels): the input patch the segment
// needs (kh rows x ((STEM_TW - 1)*stride + kw) columns x 3 channels, fp32) is staged in shared memory with coalesced row reads,
// then every thread assembles 16-byte pieces (8 columns) of the output, consecutive threads writing consecutive pieces (a row
// is kpad / 8 whole pieces, so a segment's output is one contiguous range).  History: one thread per (pixel, ky) with 2-byte
// stores ran at 0.36 TB/s (1.2 ms of the 12.5 ms batch-16 pass); 16-byte stores with gathers straight from global memory
// 0.8 TB/s (0.54 ms): 219 M scalar loads through L1.
constexpr int STEM_TW = 64;
constexpr int STEM_MAX_PATCH = 7 * ((STEM_TW - 1) * 2 + 7) * 3;     // kh <= 7, stride <= 2, kw <= 7

__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ y, int n, int h,
                                                          int w, int ho, int wo, int kh, int kw, int stride, int pad, int kpad,
                                                          int segs) {
  __shared__ float patch[STEM_MAX_PATCH];                 // [c][ky][col]
  const int seg = blockIdx.x % segs;
  const int oy = (blockIdx.x / segs) % ho;
  const int b = blockIdx.x / (segs * ho);
  const int ox0 = seg * STEM_TW;
  const int npix = min(STEM_TW, wo - ox0);
  const int pw = (npix - 1) * stride + kw;                // patch width
  const int iy0 = oy * stride - pad, ix0 = ox0 * stride - pad;
  const float* src = img + (long long)b * 3 * h * w;
  // one warp per (channel, filter row): coalesced reads of the input row segment, no index arithmetic per element
  for (int r = threadIdx.x >> 5; r < 3 * kh; r += blockDim.x >> 5) {
    const int c = r / kh, iy = iy0 + (r - c * kh);
    const float* row = src + ((long long)c * h + iy) * w;
    const bool row_ok = iy >= 0 && iy < h;
    for (int col = threadIdx.x & 31; col < pw; col += 32) {
      const int ix = ix0 + col;
      patch[r * pw + col] = (row_ok && ix >= 0 && ix < w) ? __ldg(row + ix) : 0.f;
    }
  }
  // thread = (piece q of a row, pixel lane): the eight (channel, tap) offsets of its piece are decoded ONCE
  const int pieces = kpad >> 3;
  const int lanes = blockDim.x / pieces;                  // pixels in flight per pass
  const int q = threadIdx.x % pieces, px0 = threadIdx.x / pieces;
  const int kcols = kh * kw * 3;
  int off[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int col = q * 8 + e;
    const int tap = col / 3, c = col - tap * 3;
    const int ky = tap / kw, kx = tap - ky * kw;
    off[e] = col < kcols ? (c * kh + ky) * pw + kx : -1;
  }
  __syncthreads();
  if (px0 >= lanes) return;
  __nv_bfloat16* dst = y + ((long long)(b * ho + oy) * wo + ox0) * kpad;
  for (int px = px0; px < npix; px += lanes) {
    const float* pp = patch + px * stride;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = off[e] >= 0 ? pp[off[e]] : 0.f;
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + ((long long)px * pieces + q) * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------ NHWC im2col
// y[pix, (ky*kw + kx)*c + ci] = x[n, oy*stride - pad + ky, ox*stride - pad + kx, ci]; c % 8 == 0; 16 bytes per thread.
__global__ void im2col_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h, int w,
                                   int c, int ho, int wo, int kh, int kw, int stride, int pad) {
  const int cv = c >> 3;
  const long long total = (long long)n * ho * wo * kh * kw * cv;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
    const int ci = (int)(idx % cv);
    long long r = idx / cv;
    const int kx = (int)(r % kw); r /= kw;
    const int ky = (int)(r % kh); r /= kh;
    const long long pix = r;
    const int ox = (int)(pix % wo);
    const int oy = (int)((pix / wo) % ho);
    const int b = (int)(pix / ((long long)wo * ho));
    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < h && ix >= 0 && ix < w)
      v = __ldg(reinterpret_cast<const uint4*>(x + (((long long)b * h + iy) * w + ix) * c + ci * 8));
    *reinterpret_cast<uint4*>(y + (pix * kh * kw + ky * kw + kx) * c + ci * 8) = v;
  }
}

// ------------------------------------------------------------------------------------------------ 3x3/2 max-pool (pad 1)
__global__ void maxpool_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h, int w,
                                    int c, int ho, int wo, int k, int stride, int pad) {
  const int cv = c >> 3;
  const long long total = (long long)n * ho * wo * cv;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
    const int ci = (int)(idx % cv);
    const long long pix = idx / cv;
    const int ox = (int)(pix % wo);
    const int oy = (int)((pix / wo) % ho);
    const int b = (int)(pix / ((long long)wo * ho));
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= h) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= w) continue;
        float v[8];
        ld8b(x + (((long long)b * h + iy) * w + ix) * c + ci * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
      }
    }
    *reinterpret_cast<uint4*>(y + pix * c + ci * 8) = pack8(m);
  }
}

// ------------------------------------------------------------------------------------------------ RoIPool
// torchvision.ops.roi_pool semantics (csrc/ops/cpu/roi_pool_kernel.cpp): coordinates are rounded half away from zero
// after scaling, a RoI is at least 1x1, bin borders are floor / ceil of multiples of roi_size / pooled_size, clipped to the
// map; an empty bin gives 0; otherwise the maximum (first maximum in row-major scan order for argmax).
// One thread per 16-byte piece (8 channels) of the output; ARGMAX is a compile-time switch (the extractors do not ask for it, and
// tracking eight indices per thread doubled the instruction count of this instruction-bound kernel), indices are 32-bit.
template <bool ARGMAX>
__global__ void roi_pool_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ rois,
                                     __nv_bfloat16* __restrict__ y, int* __restrict__ argmax, int num_rois, int n, int h,
                                     int w, int c, int ph, int pw, float spatial_scale) {
  const int cv = c >> 3;
  const unsigned total = (unsigned)num_rois * ph * pw * cv;
  const unsigned step = gridDim.x * blockDim.x;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
    const int ci = (int)(idx % cv);
    unsigned r = idx / cv;
    const int px = (int)(r % pw); r /= pw;
    const int py = (int)(r % ph); r /= ph;
    const int roi = (int)r;
    const float* box = rois + (long long)roi * 5;
    const int b = (int)box[0];
    const int roi_start_w = (int)roundf(box[1] * spatial_scale);
    const int roi_start_h = (int)roundf(box[2] * spatial_scale);
    const int roi_end_w = (int)roundf(box[3] * spatial_scale);
    const int roi_end_h = (int)roundf(box[4] * spatial_scale);
    const int roi_width = max(roi_end_w - roi_start_w + 1, 1);
    const int roi_height = max(roi_end_h - roi_start_h + 1, 1);
    const float bin_h = (float)roi_height / (float)ph;
    const float bin_w = (float)roi_width / (float)pw;
    int hstart = (int)floorf((float)py * bin_h);
    int wstart = (int)floorf((float)px * bin_w);
    int hend = (int)ceilf((float)(py + 1) * bin_h);
    int wend = (int)ceilf((float)(px + 1) * bin_w);
    hstart = min(max(hstart + roi_start_h, 0), h);
    hend = min(max(hend + roi_start_h, 0), h);
    wstart = min(max(wstart + roi_start_w, 0), w);
    wend = min(max(wend + roi_start_w, 0), w);
    const bool empty = (hend <= hstart) || (wend <= wstart) || b < 0 || b >= n;
    float m[8];
    int am[ARGMAX ? 8 : 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = empty ? 0.f : -FLT_MAX;
    if (ARGMAX) {
#pragma unroll
      for (int i = 0; i < 8; ++i) am[ARGMAX ? i : 0] = -1;
    }
    if (!empty) {
      const __nv_bfloat16* base = x + ((long long)b * h * w) * c + ci * 8;
      for (int iy = hstart; iy < hend; ++iy)
        for (int ix = wstart; ix < wend; ++ix) {
          float v[8];
          ld8b(base + (long long)(iy * w + ix) * c, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (ARGMAX) { if (v[i] > m[i]) { m[i] = v[i]; am[ARGMAX ? i : 0] = iy * w + ix; } }
            else m[i] = fmaxf(m[i], v[i]);
          }
        }
    }
    const long long o = (((long long)roi * ph + py) * pw + px) * c + ci * 8;
    *reinterpret_cast<uint4*>(y + o) = pack8(m);
    if (ARGMAX) {
#pragma unroll
      for (int i = 0; i < 8; ++i) argmax[o + i] = am[ARGMAX ? i : 0];
    }
  }
}

// ------------------------------------------------------------------------------------------------ RoIAlign
// torchvision.ops.roi_align semantics (csrc/ops/cpu/roi_align_kernel.cpp, roi_align_common.h), aligned = false|true:
// bilinear samples on a sampling_ratio x sampling_ratio grid per bin (adaptive ceil(roi/pooled) when ratio <= 0), average.
__device__ __forceinline__ void bilinear8(const __nv_bfloat16* __restrict__ x, int b, int h, int w, int c, int ci, float yy,
                                          float xx, float (&acc)[8]) {
  if (yy < -1.0f || yy > (float)h || xx < -1.0f || xx > (float)w) return;
  if (yy <= 0.f) yy = 0.f;
  if (xx <= 0.f) xx = 0.f;
  int y_low = (int)yy, x_low = (int)xx, y_high, x_high;
  if (y_low >= h - 1) { y_high = y_low = h - 1; yy = (float)y_low; } else { y_high = y_low + 1; }
  if (x_low >= w - 1) { x_high = x_low = w - 1; xx = (float)x_low; } else { x_high = x_low + 1; }
  const float ly = yy - (float)y_low, lx = xx - (float)x_low, hy = 1.f - ly, hx = 1.f - lx;
  const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
  float v1[8], v2[8], v3[8], v4[8];
  const __nv_bfloat16* base = x + (long long)b * h * w * c + ci * 8;
  ld8b(base + ((long long)y_low * w + x_low) * c, v1);
  ld8b(base + ((long long)y_low * w + x_high) * c, v2);
  ld8b(base + ((long long)y_high * w + x_low) * c, v3);
  ld8b(base + ((long long)y_high * w + x_high) * c, v4);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] += w1 * v1[i] + w2 * v2[i] + w3 * v3[i] + w4 * v4[i];
}

__global__ void roi_align_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ rois,
                                      __nv_bfloat16* __restrict__ y, int num_rois, int n, int h, int w, int c, int ph,
                                      int pw, float spatial_scale, int sampling_ratio, int aligned) {
  const int cv = c >> 3;
  const long long total = (long long)num_rois * ph * pw * cv;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
    const int ci = (int)(idx % cv);
    long long r = idx / cv;
    const int px = (int)(r % pw); r /= pw;
    const int py = (int)(r % ph); r /= ph;
    const int roi = (int)r;
    const float* box = rois + (long long)roi * 5;
    const int b = (int)box[0];
    const float offset = aligned ? 0.5f : 0.f;
    const float roi_start_w = box[1] * spatial_scale - offset;
    const float roi_start_h = box[2] * spatial_scale - offset;
    const float roi_end_w = box[3] * spatial_scale - offset;
    const float roi_end_h = box[4] * spatial_scale - offset;
    float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
    if (!aligned) { roi_width = fmaxf(roi_width, 1.f); roi_height = fmaxf(roi_height, 1.f); }
    const float bin_h = roi_height / (float)ph, bin_w = roi_width / (float)pw;
    const int grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_height / (float)ph);
    const int grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_width / (float)pw);
    const float count = fmaxf((float)(grid_h * grid_w), 1.f);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (b >= 0 && b < n) {
      for (int iy = 0; iy < grid_h; ++iy) {
        const float yy = roi_start_h + (float)py * bin_h + ((float)iy + 0.5f) * bin_h / (float)grid_h;
        for (int ix = 0; ix < grid_w; ++ix) {
          const float xx = roi_start_w + (float)px * bin_w + ((float)ix + 0.5f) * bin_w / (float)grid_w;
          bilinear8(x, b, h, w, c, ci, yy, xx, acc);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] /= count;
    *reinterpret_cast<uint4*>(y + (((long long)roi * ph + py) * pw + px) * c + ci * 8) = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------------ global average pool
// out[r, c] = mean_s x[r, s, c]  (AdaptiveAvgPool2d((1,1)) + flatten), fp32 output
__global__ void avgpool_nhwc_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int r, int s, int c) {
  const int cv = c >> 3;
  const long long total = (long long)r * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ci = (int)(idx % cv);
  const long long ri = idx / cv;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int si = 0; si < s; ++si) {
    float v[8];
    ld8b(x + (ri * s + si) * c + ci * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += v[i];
  }
  const float inv = 1.f / (float)s;
  float* o = out + ri * c + ci * 8;
  *reinterpret_cast<float4*>(o) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

}  // namespace vb

using namespace vb;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int vb_stem_im2col(const float* img, void* y, int32_t n, int32_t h, int32_t w, int32_t kh, int32_t kw,
                              int32_t stride, int32_t pad, int32_t kpad, void* stream) {
  VB_REQUIRE(img && y && n > 0 && h > 0 && w > 0, "null pointer or empty image");
  VB_REQUIRE(kh > 0 && kw > 0 && stride > 0 && pad >= 0 && kpad >= kh * kw * 3 && kpad % 8 == 0, "bad window / kpad");
  const int ho = (h + 2 * pad - kh) / stride + 1, wo = (w + 2 * pad - kw) / stride + 1;
  VB_REQUIRE(ho > 0 && wo > 0, "empty output");
  VB_REQUIRE(al16(y), "output must be 16-byte aligned");
  VB_REQUIRE(kh <= 7 && kw <= 7 && stride <= 2 && kpad <= 2048, "stem window above 7x7 / stride above 2 / kpad above 2048");
  const int segs = (wo + vb::STEM_TW - 1) / vb::STEM_TW;
  const long long blocks = (long long)n * ho * segs;
  VB_REQUIRE(blocks < (1ll << 31), "image batch too large");
  stem_im2col_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(img, (__nv_bfloat16*)y, n, h, w, ho, wo, kh, kw, stride,
                                                                         pad, kpad, segs);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_im2col_nhwc(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t kh, int32_t kw,
                              int32_t stride, int32_t pad, void* stream) {
  VB_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "null pointer, empty tensor or c % 8 != 0");
  VB_REQUIRE(al16(x) && al16(y), "pointers must be 16-byte aligned");
  VB_REQUIRE(kh > 0 && kw > 0 && stride > 0 && pad >= 0, "bad window");
  const int ho = (h + 2 * pad - kh) / stride + 1, wo = (w + 2 * pad - kw) / stride + 1;
  VB_REQUIRE(ho > 0 && wo > 0, "empty output");
  const long long total = (long long)n * ho * wo * kh * kw * (c / 8);
  im2col_nhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w,
                                                                              c, ho, wo, kh, kw, stride, pad);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_maxpool_nhwc(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, int32_t stride,
                               int32_t pad, void* stream) {
  VB_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "null pointer, empty tensor or c % 8 != 0");
  VB_REQUIRE(al16(x) && al16(y) && k > 0 && stride > 0 && pad >= 0 && pad < k, "alignment / window");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  VB_REQUIRE(ho > 0 && wo > 0, "empty output");
  const long long total = (long long)n * ho * wo * (c / 8);
  maxpool_nhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w,
                                                                               c, ho, wo, k, stride, pad);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_roi_pool_nhwc(const void* x, const float* rois, void* y, int32_t* argmax, int32_t num_rois, int32_t n,
                                int32_t h, int32_t w, int32_t c, int32_t ph, int32_t pw, float spatial_scale, void* stream) {
  VB_REQUIRE(x && rois && y && num_rois >= 0 && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "bad arguments");
  VB_REQUIRE(al16(x) && al16(y) && ph > 0 && pw > 0, "alignment / pooled size");
  if (num_rois == 0) return VB_OK;
  const long long total = (long long)num_rois * ph * pw * (c / 8);
  VB_REQUIRE(total < (1ll << 31), "too many RoI bins for the 32-bit index");
  if (argmax != nullptr)
    roi_pool_nhwc_kernel<true><<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rois, (__nv_bfloat16*)y,
                                                                                        argmax, num_rois, n, h, w, c, ph, pw, spatial_scale);
  else
    roi_pool_nhwc_kernel<false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rois, (__nv_bfloat16*)y,
                                                                                         nullptr, num_rois, n, h, w, c, ph, pw, spatial_scale);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_roi_align_nhwc(const void* x, const float* rois, void* y, int32_t num_rois, int32_t n, int32_t h, int32_t w,
                                 int32_t c, int32_t ph, int32_t pw, float spatial_scale, int32_t sampling_ratio, int32_t aligned,
                                 void* stream) {
  VB_REQUIRE(x && rois && y && num_rois >= 0 && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "bad arguments");
  VB_REQUIRE(al16(x) && al16(y) && ph > 0 && pw > 0, "alignment / pooled size");
  if (num_rois == 0) return VB_OK;
  const long long total = (long long)num_rois * ph * pw * (c / 8);
  roi_align_nhwc_kernel<<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rois, (__nv_bfloat16*)y,
                                                                                 num_rois, n, h, w, c, ph, pw, spatial_scale,
                                                                                 sampling_ratio, aligned);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_avgpool_nhwc(const void* x, float* out, int32_t r, int32_t s, int32_t c, void* stream) {
  VB_REQUIRE(x && out && r > 0 && s > 0 && c > 0 && c % 8 == 0 && al16(x) && al16(out), "bad arguments");
  const long long total = (long long)r * (c / 8);
  avgpool_nhwc_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, r, s, c);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}
