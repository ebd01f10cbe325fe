// Fused optimizer step over the flat parameter / gradient buffers: the reference's training loop
// (pipelines/model_training/nodes.py:795-799) runs torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0) and
// torch.optim.AdamW(lr, weight_decay=0.01).step() as hundreds of small foreach launches over 523 tensors.  Here it is two
// bandwidth-bound passes over contiguous memory:
//   1. vb_grad_sumsq : sum of squares of the gradient range -> one fp64 accumulator on the device
//   2. vb_adamw_step : clip coefficient from that accumulator, decoupled weight decay, Adam moments, parameter update and the
//                      bf16 weight shadow of the GEMM operands, all in one read-modify-write (30 bytes per parameter)
// Arithmetic follows torch's single-tensor AdamW formulas (torch/optim/adamw.py -> adam.py::_single_tensor_adam) in fp32.
#include "common.cuh"
#include "../../include/vilbert_b200.h"

namespace vb {

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ acc) {
  const long long n4 = n >> 2;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += static_cast<double>(part[i]);
    atomicAdd(acc, t);
  }
}

struct AdamParams {
  float* p; const float* g; float* m; float* v; __nv_bfloat16* shadow;
  long long n, shadow_n;
  const double* sumsq;       // may be NULL: no clipping
  float max_norm;
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt;
};

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, const AdamParams& a, float coef) {
  g *= coef;
  p *= 1.0f - a.lr * a.weight_decay;
  m = m + (g - m) * (1.0f - a.beta1);                     // exp_avg.lerp_(grad, 1 - beta1)
  v = v * a.beta2 + (1.0f - a.beta2) * g * g;             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= (a.lr / a.bc1) * (m / denom);
  return p;
}

__global__ void __launch_bounds__(256) adamw_kernel(const AdamParams a) {
  float coef = 1.0f;
  if (a.sumsq != nullptr) {
    const float total = static_cast<float>(sqrt(*a.sumsq));
    coef = fminf(a.max_norm / (total + 1e-6f), 1.0f);     // torch.nn.utils.clip_grad_norm_
  }
  const long long n4 = a.n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.g) + i);
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    adam_one(p.x, g.x, m.x, v.x, a, coef); adam_one(p.y, g.y, m.y, v.y, a, coef);
    adam_one(p.z, g.z, m.z, v.z, a, coef); adam_one(p.w, g.w, m.w, v.w, a, coef);
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
    if ((i << 2) + 3 < a.shadow_n) {
      uint2 o;
      o.x = pack_bf16x2(p.x, p.y);
      o.y = pack_bf16x2(p.z, p.w);
      reinterpret_cast<uint2*>(a.shadow)[i] = o;
    } else {
      for (int j = 0; j < 4; ++j)
        if ((i << 2) + j < a.shadow_n) a.shadow[(i << 2) + j] = __float2bfloat16_rn((&p.x)[j]);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float p = a.p[i], m = a.m[i], v = a.v[i];
    adam_one(p, a.g[i], m, v, a, coef);
    a.p[i] = p; a.m[i] = m; a.v[i] = v;
    if (i < a.shadow_n) a.shadow[i] = __float2bfloat16_rn(p);
  }
}

}  // namespace vb

extern "C" int vb_grad_sumsq(const float* grad, int64_t n, double* acc, void* stream) {
  VB_REQUIRE(grad && acc && n >= 0, "null pointer");
  VB_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "grad must be 16-byte aligned");
  if (n == 0) return VB_OK;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  vb::grad_sumsq_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(grad, n, acc);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}

extern "C" int vb_adamw_step(const vb_adamw_args* args, void* stream) {
  VB_REQUIRE(args != nullptr, "null args");
  const vb_adamw_args& x = *args;
  VB_REQUIRE(x.param && x.grad && x.exp_avg && x.exp_avg_sq && x.n >= 0, "null pointer");
  VB_REQUIRE(((reinterpret_cast<uintptr_t>(x.param) | reinterpret_cast<uintptr_t>(x.grad) | reinterpret_cast<uintptr_t>(x.exp_avg) |
               reinterpret_cast<uintptr_t>(x.exp_avg_sq)) & 15) == 0, "buffers must be 16-byte aligned");
  VB_REQUIRE(x.shadow_n == 0 || (x.shadow != nullptr && (reinterpret_cast<uintptr_t>(x.shadow) & 7) == 0 && x.shadow_n <= x.n),
             "shadow missing / misaligned / longer than the range");
  VB_REQUIRE(x.step >= 1 && x.beta1 >= 0.f && x.beta1 < 1.f && x.beta2 >= 0.f && x.beta2 < 1.f, "bad hyper-parameters");
  if (x.n == 0) return VB_OK;
  vb::AdamParams a;
  a.p = x.param; a.g = x.grad; a.m = x.exp_avg; a.v = x.exp_avg_sq; a.shadow = static_cast<__nv_bfloat16*>(x.shadow);
  a.n = x.n; a.shadow_n = x.shadow_n; a.sumsq = x.max_norm > 0.f ? x.grad_sumsq : nullptr; a.max_norm = x.max_norm;
  a.lr = x.lr; a.beta1 = x.beta1; a.beta2 = x.beta2; a.eps = x.eps; a.weight_decay = x.weight_decay;
  a.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(x.beta1), static_cast<double>(x.step)));
  a.bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(x.beta2), static_cast<double>(x.step))));
  VB_REQUIRE(a.sumsq != nullptr || x.max_norm <= 0.f, "max_norm > 0 needs grad_sumsq");
  long long blocks = (x.n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  vb::adamw_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  VB_CUDA_CHECK(cudaGetLastError());
  return VB_OK;
}
