// Feature-store ingest (reference pipelines/data_processing/lmdb_dataset.py:143-208): one batch of
// detectron.lmdb records arrives in HBM as raw fp32 region features [rows, F] and raw fp32 boxes [rows, box_stride >= 4];
// one launch produces what the encoder consumes: the features rounded to bf16 (the A operand of the image-embedding GEMM)
// and the 5-wide spatial rows [x1/1000, y1/1000, x2/1000, y2/1000, ((x2-x1)*(y2-y1))/1e6].  The box arithmetic is the
// reference's numpy float32 sequence with explicit round-to-nearest intrinsics (no FMA contraction, IEEE division), so the
// spatial rows are bit-equal to `_process_boxes`.  HBM-bound: 4 B read + 2 B written per feature element, each touched once.
#include "common.cuh"
#include "../../include/vilbert_b200.h"

namespace vb {

constexpr int INGEST_THREADS = 256;

__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// blocks [0, cast_blocks): grid-stride cast, 8 elements per thread per trip; blocks [cast_blocks, gridDim.x): one box row
// per thread.
__global__ void __launch_bounds__(INGEST_THREADS)
lmdb_regions_kernel(const float* __restrict__ feat, long long n_feat, __nv_bfloat16* __restrict__ feat_out,
                    const float* __restrict__ boxes, int rows, int box_stride, float box_div, float area_div,
                    float* __restrict__ spatial, int cast_blocks) {
  if ((int)blockIdx.x < cast_blocks) {
    const long long stride = (long long)cast_blocks * INGEST_THREADS * 8;
    for (long long i = ((long long)blockIdx.x * INGEST_THREADS + threadIdx.x) * 8; i < n_feat; i += stride) {
      const float4 a = ld_stream_f4(feat + i), b = ld_stream_f4(feat + i + 4);
      uint4 u;
      u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
      *reinterpret_cast<uint4*>(feat_out + i) = u;
    }
    return;
  }
  const int r = ((int)blockIdx.x - cast_blocks) * INGEST_THREADS + threadIdx.x;
  if (r >= rows) return;
  const float* b = boxes + (long long)r * box_stride;
  const float x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
  const float w = __fsub_rn(x2, x1), h = __fsub_rn(y2, y1);                 // lmdb_dataset.py:194-195
  float* o = spatial + (long long)r * 5;
  o[0] = __fdiv_rn(x1, box_div);                                            // :200-203
  o[1] = __fdiv_rn(y1, box_div);
  o[2] = __fdiv_rn(x2, box_div);
  o[3] = __fdiv_rn(y2, box_div);
  o[4] = __fdiv_rn(__fmul_rn(w, h), area_div);                              // :196
}

}  // namespace vb

extern "C" int vb_lmdb_regions(const float* features, void* features_bf16, int64_t n_features, const float* boxes,
                               float* spatial, int32_t rows, int32_t box_stride, float box_div, float area_div, void* stream) {
  VB_REQUIRE(n_features >= 0 && rows >= 0, "negative size");
  VB_REQUIRE(n_features == 0 || (features && features_bf16), "null feature pointer");
  VB_REQUIRE(n_features % 8 == 0, "feature element count must be a multiple of 8");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(features) & 15) == 0 && (reinterpret_cast<uintptr_t>(features_bf16) & 15) == 0,
             "feature buffers must be 16-byte aligned");
  VB_REQUIRE(rows == 0 || (boxes && spatial && box_stride >= 4), "boxes need at least four columns");
  long long cast_blocks = (n_features / 8 + vb::INGEST_THREADS - 1) / vb::INGEST_THREADS;
  if (cast_blocks > 148 * 8) cast_blocks = 148 * 8;          // eight resident CTAs per SM, grid-stride beyond that
  const int box_blocks = (rows + vb::INGEST_THREADS - 1) / vb::INGEST_THREADS;
  if (cast_blocks + box_blocks == 0) return VB_OK;
  vb::lmdb_regions_kernel<<<(unsigned)(cast_blocks + box_blocks), vb::INGEST_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      features, n_features, static_cast<__nv_bfloat16*>(features_bf16), boxes, rows, box_stride, box_div, area_div, spatial,
      (int)cast_blocks);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}
