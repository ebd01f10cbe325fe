// Region Proposal Network post-processing of the Visual Genome Faster R-CNN extractor (reference
// models/feature_extractors/fasterrcnn_vg_rpn.py:60-174, 442-469) on the device, without a host read between the steps:
//
//   rpn_decode      objectness softmax, anchors (12 per cell), box deltas, clipping, min-size filter   :85-104, 106-174, 444-450
//   rank_sort_desc  stable descending order of a score vector (rank counting, many CTAs)               torch.topk :455-458, :428
//   gather_sorted   the first min(cap, #valid) boxes / scores in that order                            :455-458
//   nms_sorted      greedy NMS over an already sorted list, stopping at `max_keep` survivors           :461-465
//
// The convolutions of the RPN head are vb_gemm_bf16 (3x3 as implicit GEMM); everything here is fp32 arithmetic in the reference's
// operation order with explicit round-to-nearest intrinsics (no FMA contraction); exp is expf (differs from torch's CPU exp in the
// last place: proposals agree to ~1e-7 relative, tests/test_fasterrcnn_vg_rpn.py).
#include <cfloat>

#include "common.cuh"
#include "../../include/vilbert_b200.h"

namespace vb {

struct RpnAnchors { float a[16][4]; };

// heads: fp32 [cells, ld]: columns [0, 2A) objectness logits (anchor k: columns 2k, 2k+1 = background, foreground), columns
// [2A, 6A) box deltas (anchor k: 4 columns from 2A + 4k).  One thread per (cell, anchor), anchor fastest -- the reference's order.
__global__ void rpn_decode_kernel(const float* __restrict__ heads, int ld, int fh, int fw, int num_anchors, RpnAnchors base,
                                  float stride, float img_h, float img_w, float min_size, float* __restrict__ boxes,
                                  float* __restrict__ scores, int* __restrict__ num_valid) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = fh * fw * num_anchors;
  int ok = 0;
  if (idx < total) {
    const int cell = idx / num_anchors, k = idx - cell * num_anchors;
    const int iy = cell / fw, ix = cell - iy * fw;
    const float* row = heads + static_cast<long long>(cell) * ld;
    const float l0 = row[2 * k], l1 = row[2 * k + 1];
    const float m = fmaxf(l0, l1);
    const float e0 = expf(__fsub_rn(l0, m)), e1 = expf(__fsub_rn(l1, m));
    const float fg = __fdiv_rn(e1, __fadd_rn(e0, e1));
    const float* d = row + 2 * num_anchors + 4 * k;
    const float sx = static_cast<float>(ix) * stride + floorf(stride * 0.5f), sy = static_cast<float>(iy) * stride + floorf(stride * 0.5f);
    const float x1 = __fadd_rn(base.a[k][0], sx), y1 = __fadd_rn(base.a[k][1], sy);
    const float x2 = __fadd_rn(base.a[k][2], sx), y2 = __fadd_rn(base.a[k][3], sy);
    const float w = __fsub_rn(x2, x1), h = __fsub_rn(y2, y1);
    const float cx = __fadd_rn(x1, __fmul_rn(0.5f, w)), cy = __fadd_rn(y1, __fmul_rn(0.5f, h));
    const float dw = fminf(d[2], 4.0f), dh = fminf(d[3], 4.0f);
    const float pcx = __fadd_rn(__fmul_rn(d[0], w), cx), pcy = __fadd_rn(__fmul_rn(d[1], h), cy);
    const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
    float4 b;
    b.x = fminf(fmaxf(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), 0.f), img_w);
    b.y = fminf(fmaxf(__fsub_rn(pcy, __fmul_rn(0.5f, ph)), 0.f), img_h);
    b.z = fminf(fmaxf(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 0.f), img_w);
    b.w = fminf(fmaxf(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 0.f), img_h);
    *reinterpret_cast<float4*>(boxes + 4 * static_cast<long long>(idx)) = b;
    ok = (__fsub_rn(b.z, b.x) >= min_size && __fsub_rn(b.w, b.y) >= min_size) ? 1 : 0;
    scores[idx] = ok ? fg : -INFINITY;            // filtered proposals sort behind every real one
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, ok != 0);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(num_valid, __popc(ballot));
}

// order[rank(i)] = i with rank(i) = #{j : s_j > s_i or (s_j == s_i and j < i)}: torch.sort(descending, stable).  Elements at or
// beyond *limit (optional) count as -inf.  One thread per element, the scores streamed through shared memory in tiles.
constexpr int RANK_TILE = 2048;
__global__ void __launch_bounds__(256) rank_sort_desc_kernel(const float* __restrict__ scores, int n, const int* __restrict__ limit,
                                                             int* __restrict__ order) {
  __shared__ float tile[RANK_TILE];
  const int lim = limit ? min(*limit, n) : n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float si = (i < lim) ? scores[i] : -INFINITY;
  int rank = 0;
  for (int t0 = 0; t0 < n; t0 += RANK_TILE) {
    const int cnt = min(RANK_TILE, n - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) tile[j] = (t0 + j < lim) ? scores[t0 + j] : -INFINITY;
    __syncthreads();
    if (i < n) {
      for (int j = 0; j < cnt; ++j) {
        const float sj = tile[j];
        rank += (sj > si || (sj == si && t0 + j < i)) ? 1 : 0;
      }
    }
  }
  if (i < n) order[rank] = i;
}

// out[r] = in[order[r]] for r < min(cap, *num_valid); *count = that minimum
__global__ void gather_sorted_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int* __restrict__ order,
                                     const int* __restrict__ num_valid, int cap, float* __restrict__ out_boxes,
                                     float* __restrict__ out_scores, int* __restrict__ count) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(cap, *num_valid);
  if (r == 0) *count = n;
  if (r >= cap) return;
  if (r < n) {
    const int j = order[r];
    *reinterpret_cast<float4*>(out_boxes + 4 * r) = *reinterpret_cast<const float4*>(boxes + 4 * static_cast<long long>(j));
    out_scores[r] = scores[j];
  } else {
    *reinterpret_cast<float4*>(out_boxes + 4 * r) = make_float4(0.f, 0.f, 0.f, 0.f);
    out_scores[r] = -INFINITY;
  }
}

// Greedy NMS over boxes ALREADY in descending score order (torchvision.ops.nms semantics, CPU kernel: fp32 IoU against a double
// threshold); keep[0 .. *num_keep) = positions of the survivors.  The reference keeps only the first `max_keep` survivors, and a
// greedy decision never depends on later boxes, so the scan stops there.  One CTA; `suppressed` lives in shared memory.
constexpr int NMS_SORTED_MAX = 8192;
__global__ void __launch_bounds__(1024, 1)
nms_sorted_kernel(const float* __restrict__ boxes, const int* __restrict__ count, double iou_threshold, int max_keep,
                  int* __restrict__ keep, int* __restrict__ num_keep) {
  __shared__ unsigned char suppressed[NMS_SORTED_MAX];
  __shared__ int s_kept;
  const int n = min(*count, NMS_SORTED_MAX);
  for (int i = threadIdx.x; i < n; i += blockDim.x) suppressed[i] = 0;
  if (threadIdx.x == 0) s_kept = 0;
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    if (suppressed[i]) continue;                 // uniform: written before the barrier below
    if (s_kept >= max_keep) break;               // uniform as well
    __syncthreads();                             // everyone has read s_kept / suppressed[i]
    if (threadIdx.x == 0) keep[s_kept++] = i;
    const float4 bi = *reinterpret_cast<const float4*>(boxes + 4 * i);
    const float iarea = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
      if (suppressed[j]) continue;
      const float4 bj = *reinterpret_cast<const float4*>(boxes + 4 * j);
      const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
      const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
      const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1));
      const float h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
      const float inter = __fmul_rn(w, h);
      const float jarea = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
      if (static_cast<double>(ovr) > iou_threshold) suppressed[j] = 1;
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) *num_keep = s_kept;
}

}  // namespace vb

extern "C" int vb_rpn_decode(const float* heads, int32_t ld, int32_t fh, int32_t fw, int32_t num_anchors, const float* base_anchors,
                             float stride, float img_h, float img_w, float min_size, float* boxes, float* scores, int32_t* num_valid,
                             void* stream) {
  VB_REQUIRE(heads && base_anchors && boxes && scores && num_valid, "null pointer");
  VB_REQUIRE(fh > 0 && fw > 0 && num_anchors > 0 && num_anchors <= 16 && ld >= 6 * num_anchors, "bad geometry (at most 16 anchors per cell)");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "boxes must be 16-byte aligned");
  vb::RpnAnchors base;
  for (int k = 0; k < num_anchors; ++k)
    for (int c = 0; c < 4; ++c) base.a[k][c] = base_anchors[4 * k + c];      // HOST array
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VB_CUDA_CHECK(cudaMemsetAsync(num_valid, 0, sizeof(int32_t), s));
  const int total = fh * fw * num_anchors;
  vb::rpn_decode_kernel<<<(total + 255) / 256, 256, 0, s>>>(heads, ld, fh, fw, num_anchors, base, stride, img_h, img_w, min_size, boxes,
                                                           scores, num_valid);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_rank_sort_desc(const float* scores, int32_t n, const int32_t* limit, int32_t* order, void* stream) {
  VB_REQUIRE(scores && order && n >= 0, "null pointer");
  if (n == 0) return VB_OK;
  vb::rank_sort_desc_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, n, limit, order);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_gather_sorted(const float* boxes, const float* scores, const int32_t* order, const int32_t* num_valid, int32_t cap,
                                float* out_boxes, float* out_scores, int32_t* count, void* stream) {
  VB_REQUIRE(boxes && scores && order && num_valid && out_boxes && out_scores && count && cap > 0, "null pointer");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_boxes) & 15) == 0, "boxes must be 16-byte aligned");
  vb::gather_sorted_kernel<<<(cap + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes, scores, order, num_valid, cap, out_boxes,
                                                                                           out_scores, count);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_nms_sorted(const float* boxes, const int32_t* count, double iou_threshold, int32_t max_keep, int32_t* keep,
                             int32_t* num_keep, void* stream) {
  VB_REQUIRE(boxes && count && keep && num_keep && max_keep > 0, "null pointer");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "boxes must be 16-byte aligned");
  vb::nms_sorted_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(boxes, count, iou_threshold, max_keep, keep, num_keep);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}
