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
nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int n, float iou_threshold,
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
      if (ovr > iou_threshold) suppressed[j] = 1;
    }
    __syncthreads();
  }
  if (tid == 0) *num_keep = s_count;
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

extern "C" int vb_nms(const float* boxes, const float* scores, int32_t n, float iou_threshold, int32_t* workspace,
                      int32_t* keep, int32_t* num_keep, void* stream) {
  VB_REQUIRE(boxes && scores && workspace && keep && num_keep && n >= 0, "null pointer");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "boxes must be 16-byte aligned");
  vb::nms_kernel<<<1, vb::NMS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(boxes, scores, n, iou_threshold, workspace,
                                                                               workspace + n, keep, num_keep);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}
