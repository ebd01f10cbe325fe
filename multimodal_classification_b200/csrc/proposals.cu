// Region-proposal selection of the ResNet-152 RoI extractor (reference models/feature_extractors/resnet152_roi.py:251-293):
// the sliding-window candidates are scored by "medium area" and thinned with torchvision.ops.nms.  Which 36 boxes survive is
// decided by exact float comparisons between many tied scores, so everything here is bit-exact fp32 arithmetic in the
// reference's operation order (explicit round-to-nearest intrinsics: no FMA contraction), a STABLE descending sort, and the
// greedy suppression order of torchvision's CPU kernel (csrc/ops/cpu/nms_kernel.cpp).  One CTA; runs once per image size.
#include "common.cuh"
#include "../../include/vilbert_b200.h"

namespace vb {

constexpr int NMS_THREADS = 1024;

// score[i] = 1 - |((x2 - x1) / W) * ((y2 - y1) / H) - 0.15|      (resnet152_roi.py:262-270, fp32 tensor ops)
__global__ void box_area_score_kernel(const float* __restrict__ boxes, int n, float img_w, float img_h, float target,
                                      float* __restrict__ scores) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 b = *reinterpret_cast<const float4*>(boxes + 4 * i);
  const float w = __fdiv_rn(__fsub_rn(b.z, b.x), img_w);
  const float h = __fdiv_rn(__fsub_rn(b.w, b.y), img_h);
  const float area = __fmul_rn(w, h);
  scores[i] = __fsub_rn(1.0f, fabsf(__fsub_rn(area, target)));
}

// keep[0 .. *num_keep) = indices of the surviving boxes in descending-score order.  order / suppressed: workspace [n].
__global__ void __launch_bounds__(NMS_THREADS, 1)
nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int n, double iou_threshold,
           int* __restrict__ order, int* __restrict__ suppressed, int* __restrict__ keep, int* __restrict__ num_keep) {
  const int tid = threadIdx.x;
  // stable descending sort by rank counting: ties keep their original order, as torch.sort(stable=True, descending=True)
  for (int i = tid; i < n; i += NMS_THREADS) {
    const float si = scores[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float sj = scores[j];
      rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
    }
    order[rank] = i;
    suppressed[i] = 0;
  }
  __shared__ int s_count;
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int oi = 0; oi < n; ++oi) {
    const int i = order[oi];
    if (suppressed[i] != 0) continue;          // uniform: every thread reads the same flag after the barrier below
    if (tid == 0) keep[s_count++] = i;
    const float4 bi = *reinterpret_cast<const float4*>(boxes + 4 * i);
    const float iarea = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
    for (int oj = oi + 1 + tid; oj < n; oj += NMS_THREADS) {
      const int j = order[oj];
      const float4 bj = *reinterpret_cast<const float4*>(boxes + 4 * j);
      const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
      const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
      const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1));
      const float h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
      const float inter = __fmul_rn(w, h);
      const float jarea = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
      // torchvision's CPU kernel holds the threshold as a double: the fp32 IoU is widened for the comparison, which
      // differs from an fp32 compare exactly when IoU == float(threshold) and the threshold (0.3) is not representable
      if (static_cast<double>(ovr) > iou_threshold) suppressed[j] = 1;
    }
    __syncthreads();
  }
  if (tid == 0) *num_keep = s_count;
}

// out[r] = max_{c0 <= c < c1} x[r, c]: the proposal score of the Visual Genome extractor (fasterrcnn_vg.py:360-363,
// ``cls_scores[:, 1:].max(dim=1)``).  One warp per row; fp32 max is exact, so the order of the comparisons does not matter.
__global__ void rowmax_kernel(const float* __restrict__ x, int rows, int ld, int c0, int c1, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = x + static_cast<size_t>(row) * ld;
  float m = -INFINITY;
  for (int c = c0 + (threadIdx.x & 31); c < c1; c += 32) m = fmaxf(m, p[c]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) out[row] = m;
}

// Region selection of the Visual Genome extractor (fasterrcnn_vg.py:367-411, 436-469) without a host round trip: region r of
// the output is candidate keep[min(r, *num_keep - 1)] (``nms`` leaves the survivors in descending score order, so the
// reference's top-k over them is their first ``regions`` entries; ``_pad_regions`` repeats the last one).  Block r copies the
// candidate's box, its normalised [x1/W, y1/H, x2/W, y2/H, area] row (``_normalize_boxes``: fp32, one rounding per
// operation, clamp to [0, 1]) and, when ``feat_src`` is given, its 2048-wide feature row.
__global__ void select_regions_kernel(const float* __restrict__ cand, const int* __restrict__ keep,
                                      const int* __restrict__ num_keep, float img_w, float img_h,
                                      const float* __restrict__ feat_src, int feat_dim, float* __restrict__ boxes,
                                      float* __restrict__ spatial, float* __restrict__ feat_dst, int* __restrict__ index,
                                      float* __restrict__ rois, float batch_index, float box_div) {
  const int r = blockIdx.x;
  const int nk = *num_keep;
  if (nk <= 0) return;
  const int j = keep[min(r, nk - 1)];
  if (threadIdx.x == 0) {
    const float4 b = *reinterpret_cast<const float4*>(cand + 4 * j);
    if (boxes) *reinterpret_cast<float4*>(boxes + 4 * r) = b;
    if (index) index[r] = j;
    if (rois) {                                   // (batch index, x1, y1, x2, y2): the RoIPool operand of the chosen boxes
      float* q = rois + 5 * r;
      q[0] = batch_index; q[1] = b.x; q[2] = b.y; q[3] = b.z; q[4] = b.w;
    }
    if (spatial) {
      const float x1 = fminf(fmaxf(__fdiv_rn(__fdiv_rn(b.x, box_div), img_w), 0.0f), 1.0f);
      const float y1 = fminf(fmaxf(__fdiv_rn(__fdiv_rn(b.y, box_div), img_h), 0.0f), 1.0f);
      const float x2 = fminf(fmaxf(__fdiv_rn(__fdiv_rn(b.z, box_div), img_w), 0.0f), 1.0f);
      const float y2 = fminf(fmaxf(__fdiv_rn(__fdiv_rn(b.w, box_div), img_h), 0.0f), 1.0f);
      float* s = spatial + 5 * r;
      s[0] = x1; s[1] = y1; s[2] = x2; s[3] = y2;
      s[4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    }
  }
  if (feat_src) {
    const float4* src = reinterpret_cast<const float4*>(feat_src + static_cast<size_t>(j) * feat_dim);
    float4* dst = reinterpret_cast<float4*>(feat_dst + static_cast<size_t>(r) * feat_dim);
    for (int c = threadIdx.x; c < feat_dim / 4; c += blockDim.x) dst[c] = src[c];
  }
}

}  // namespace vb

extern "C" int vb_box_area_score(const float* boxes, int32_t n, float img_w, float img_h, float target_area, float* scores,
                                 void* stream) {
  VB_REQUIRE(boxes && scores && n >= 0, "null pointer");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "boxes must be 16-byte aligned");
  if (n == 0) return VB_OK;
  vb::box_area_score_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes, n, img_w, img_h, target_area,
                                                                                         scores);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_nms(const float* boxes, const float* scores, int32_t n, double iou_threshold, int32_t* workspace,
                      int32_t* keep, int32_t* num_keep, void* stream) {
  VB_REQUIRE(boxes && scores && workspace && keep && num_keep && n >= 0, "null pointer");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "boxes must be 16-byte aligned");
  vb::nms_kernel<<<1, vb::NMS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(boxes, scores, n, iou_threshold, workspace,
                                                                               workspace + n, keep, num_keep);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_rowmax_f32(const float* x, int32_t rows, int32_t ld, int32_t col_begin, int32_t col_end, float* out,
                             void* stream) {
  VB_REQUIRE(x && out && rows >= 0, "null pointer");
  VB_REQUIRE(0 <= col_begin && col_begin < col_end && col_end <= ld, "column range must lie inside a row");
  if (rows == 0) return VB_OK;
  vb::rowmax_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, ld, col_begin, col_end, out);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_select_regions(const float* candidates, const int32_t* keep, const int32_t* num_keep, int32_t regions,
                                 float img_w, float img_h, const float* feat_src, int32_t feat_dim, float* boxes,
                                 float* spatial, float* feat_dst, int32_t* index, float* rois, float batch_index,
                                 float box_div, void* stream) {
  VB_REQUIRE(box_div > 0.f, "box_div must be positive");
  VB_REQUIRE(candidates && keep && num_keep && regions >= 0, "null pointer");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(candidates) & 15) == 0 && (reinterpret_cast<uintptr_t>(boxes) & 15) == 0,
             "boxes must be 16-byte aligned");
  VB_REQUIRE((feat_src == nullptr) == (feat_dst == nullptr), "feature source and destination go together");
  VB_REQUIRE(!feat_src || (feat_dim > 0 && feat_dim % 4 == 0 && (reinterpret_cast<uintptr_t>(feat_src) & 15) == 0 &&
                           (reinterpret_cast<uintptr_t>(feat_dst) & 15) == 0),
             "feature rows must be 16-byte aligned and a multiple of 4 wide");
  if (regions == 0) return VB_OK;
  vb::select_regions_kernel<<<regions, 128, 0, static_cast<cudaStream_t>(stream)>>>(candidates, keep, num_keep, img_w, img_h,
                                                                                   feat_src, feat_dim, boxes, spatial,
                                                                                   feat_dst, index, rois, batch_index, box_div);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}
