#include "tensormap.h"

#include <stdio.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace vb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  uint64_t v[8];
  bool operator==(const MapKey& o) const {
    for (int i = 0; i < 8; ++i)
      if (v[i] != o.v[i]) return false;
    return true;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 8; ++i) {
      h ^= k.v[i];
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

static std::mutex g_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_cache;

static int encode_cached(CUtensorMap* out, const MapKey& key, uint32_t rank, const void* base, const cuuint64_t* dims,
                         const cuuint64_t* strides_bytes, const cuuint32_t* box,
                         CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                         CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
      *out = it->second;
      return VB_OK;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    vb_set_last_error("cuTensorMapEncodeTiled", "CUDA driver entry point not available (no GPU driver?)");
    return VB_ERR_NO_DRIVER;
  }
  const cuuint32_t elem_strides[3] = {1, 1, 1};
  CUresult r = fn(out, dtype, rank, const_cast<void*>(base), dims, strides_bytes, box,
                  elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static thread_local char buf[256];
    snprintf(buf, sizeof(buf), "CUresult %d (rank %u dims %llu,%llu,%llu box %u,%u,%u stride %llu)", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull,
             box[0], box[1], rank > 2 ? box[2] : 0u, (unsigned long long)strides_bytes[0]);
    vb_set_last_error("cuTensorMapEncodeTiled", buf);
    return VB_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_cache.size() > 65536) g_cache.clear();
  g_cache.emplace(key, *out);
  return VB_OK;
}

int make_tensor_map_2d(CUtensorMap* out, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                       int box_outer) {
  MapKey key = {{reinterpret_cast<uint64_t>(base), (uint64_t)inner, (uint64_t)outer, (uint64_t)ld,
                 (uint64_t)box_inner, (uint64_t)box_outer, 2, 0}};
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  return encode_cached(out, key, 2, base, dims, strides, box);
}

int make_tensor_map_2d_f32(CUtensorMap* out, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                           int box_outer) {
  MapKey key = {{reinterpret_cast<uint64_t>(base), (uint64_t)inner, (uint64_t)outer, (uint64_t)ld,
                 (uint64_t)box_inner, (uint64_t)box_outer, 2, /*fp32*/ 4}};
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  return encode_cached(out, key, 2, base, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

int make_tensor_map_2d_sw64(CUtensorMap* out, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                            int box_outer) {
  MapKey key = {{reinterpret_cast<uint64_t>(base), (uint64_t)inner, (uint64_t)outer, (uint64_t)ld,
                 (uint64_t)box_inner, (uint64_t)box_outer, 2, /*bf16, 64-byte swizzle*/ 64}};
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  return encode_cached(out, key, 2, base, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
}

int make_tensor_map_3d(CUtensorMap* out, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t s1, int64_t s2,
                       int box0, int box1, int box2) {
  MapKey key = {{reinterpret_cast<uint64_t>(base), (uint64_t)d0, (uint64_t)d1, (uint64_t)d2,
                 (uint64_t)s1 ^ ((uint64_t)s2 << 32), (uint64_t)box0 | ((uint64_t)box1 << 16) | ((uint64_t)box2 << 32),
                 3, (uint64_t)s2}};
  const cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  const cuuint64_t strides[2] = {(cuuint64_t)s1 * 2, (cuuint64_t)s2 * 2};
  const cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, (cuuint32_t)box2};
  return encode_cached(out, key, 3, base, dims, strides, box);
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  });
  return fn;
}

int make_tensor_map_im2col(CUtensorMap* out, const void* base, int n, int h, int w, int c, int kh, int kw, int stride,
                           int pad, int pixels) {
  MapKey key = {{reinterpret_cast<uint64_t>(base), ((uint64_t)n << 32) | (uint64_t)c, ((uint64_t)h << 32) | (uint64_t)w,
                 ((uint64_t)kh << 48) | ((uint64_t)kw << 32) | ((uint64_t)stride << 16) | (uint64_t)pad, (uint64_t)pixels,
                 /*im2col*/ 0x696d32636f6cull, 4, 0}};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
      *out = it->second;
      return VB_OK;
    }
  }
  EncodeIm2colFn fn = get_encode_im2col_fn();
  if (fn == nullptr) {
    vb_set_last_error("cuTensorMapEncodeIm2col", "CUDA driver entry point not available (no GPU driver?)");
    return VB_ERR_NO_DRIVER;
  }
  // (C, W, H, N), strides in bytes of W, H, N.  The bounding box of the filter's top-left corner: it starts `pad` pixels
  // before the picture and ends where the LAST tap still has to fit, i.e. upper corner = pad - (k - 1) relative to the
  // picture's end; the base pixel of output (oy, ox) is (oy * stride - pad, ox * stride - pad), the tap is added by the
  // instruction's offsets, and the traversal stride between consecutive output pixels is the convolution stride.
  const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  const int lower[2] = {-pad, -pad};
  const int upper[2] = {pad - (kw - 1), pad - (kh - 1)};
  const cuuint32_t traversal[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper,
                  /*channelsPerPixel*/ 64, /*pixelsPerColumn*/ (cuuint32_t)pixels, traversal, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static thread_local char buf[256];
    snprintf(buf, sizeof(buf), "CUresult %d (nhwc %d,%d,%d,%d window %dx%d stride %d pad %d pixels %d)", (int)r, n, h, w, c, kh,
             kw, stride, pad, pixels);
    vb_set_last_error("cuTensorMapEncodeIm2col", buf);
    return VB_ERR_CUDA;
  }
  // Drivers up to CUDA 13.1 set a descriptor bit for tensors below 128 KB that makes im2col loads of such tensors fault; the
  // CUTLASS im2col descriptor factory clears it the same way (cute/atom/copy_traits_sm90_im2col.hpp).
  int driver = 0;
  if (cudaDriverGetVersion(&driver) == cudaSuccess && driver <= 13010 && (uint64_t)n * h * w * c * 2 < 131072ull)
    reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_cache.size() > 65536) g_cache.clear();
  g_cache.emplace(key, *out);
  return VB_OK;
}

}  // namespace vb
