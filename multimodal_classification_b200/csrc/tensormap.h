// Host-side TMA descriptor factory.  cuTensorMapEncodeTiled is resolved through
// cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda.so (it must load —
// though not compute — on a CPU-only box for the ABI tests).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace vb {
// bf16, row-major [outer, inner] with `ld` elements between rows, 128-byte swizzle, OOB reads give zeros.
int make_tensor_map_2d(CUtensorMap* out, const void* base, int64_t inner, int64_t outer, int64_t ld,
                       int box_inner, int box_outer);
// fp32 variant (epilogue stores / reduce-adds of fp32 gradients); box_inner * 4 bytes must be <= 128.
int make_tensor_map_2d_f32(CUtensorMap* out, const void* base, int64_t inner, int64_t outer, int64_t ld,
                           int box_inner, int box_outer);
// bf16 with the 64-byte swizzle (epilogue boxes of 32 columns = 64-byte rows); box_inner * 2 bytes must be <= 64.
int make_tensor_map_2d_sw64(CUtensorMap* out, const void* base, int64_t inner, int64_t outer, int64_t ld,
                            int box_inner, int box_outer);
// bf16, [d2, d1, d0] with element strides (s2, s1, 1); 128-byte swizzle.
int make_tensor_map_3d(CUtensorMap* out, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t s1,
                       int64_t s2, int box0, int box1, int box2);
// bf16 NHWC activation [n, h, w, c] read in IM2COL mode (cuTensorMapEncodeIm2col): one load fetches `pixels` consecutive output
// pixels x 64 channels of ONE filter tap (the tap is the instruction's {kx, ky} offset operand) of a kh x kw / stride / pad
// convolution, laid out in shared memory exactly like a K-major [pixels, 64] box with the 128-byte swizzle; input positions
// outside the picture read as zeros.  c % 64 == 0.
int make_tensor_map_im2col(CUtensorMap* out, const void* base, int n, int h, int w, int c, int kh, int kw, int stride,
                           int pad, int pixels);
}  // namespace vb
