/*
 * vilbert_b200.h — C ABI of libvilbert_b200.so, the sm_100a kernel library behind the ViLBERT hot path
 * (two-stream encoder forward/backward + ResNet-152 RoI feature stage).
 *
 * Boundary rules:
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross the ABI.
 *   - every pointer is a DEVICE pointer owned by the caller unless a comment says "host";
 *     the library never allocates or frees caller-visible memory; workspaces are passed in.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - return 0 on success, a negative vb_status otherwise; never throws.  vb_last_error() returns a
 *     thread-local, human readable message for the last failing call.
 *   - bf16 tensors are raw uint16 storage (`__nv_bfloat16`), row-major, leading dimension in ELEMENTS.
 *
 * Each entry point cites the reference call site it replaces (paths relative to
 * /root/reference/src/multimodalclassification/).
 */
#ifndef VILBERT_B200_H
#define VILBERT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vb_status {
  VB_OK = 0,
  VB_ERR_BAD_ARG = -1,
  VB_ERR_CUDA = -2,
  VB_ERR_UNSUPPORTED = -3,
  VB_ERR_NO_DRIVER = -4
} vb_status;

/* library info ---------------------------------------------------------------------------------- */
int vb_abi_version(void);              /* bumped whenever a signature changes */
const char* vb_last_error(void);       /* host string, valid until the next failing call on this thread */
const char* vb_build_info(void);       /* "sm_100a; nvcc x.y; <date>" */

/* ------------------------------------------------------------------------------------------------
 * Dense contraction on tcgen05 / TMEM, TMA-fed.
 *   D[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
 * A is given either K-major (row-major [M,K], `a_mn_major=0`) or MN-major (row-major [K,M]);
 * likewise B ([N,K] or [K,N]).  This one kernel serves
 *   forward  nn.Linear           models/vilbert_facebook_arch.py:127-129,157,184,198,262-269,329,334,101
 *            A = X[M,K], B = W[N,K]                      (both K-major)
 *   dgrad    dX = dY W           autograd of the above;   A = dY[M,N'] K-major, B = W[N',K'] MN-major
 *   wgrad    dW = dY^T X         autograd of the above;   A = dY MN-major, B = X MN-major, fp32 out
 *   1x1 / im2col convolutions of torchvision ResNet-152  models/feature_extractors/resnet152_roi.py:49-74
 * ---------------------------------------------------------------------------------------------- */
typedef enum vb_act { VB_ACT_NONE = 0, VB_ACT_GELU = 1, VB_ACT_RELU = 2, VB_ACT_TANH = 3 } vb_act;
typedef enum vb_aux_mode {
  VB_AUX_NONE = 0,
  VB_AUX_ADD = 1,          /* v += aux[m,n]            (residual, before the activation) */
  VB_AUX_MUL_GELU_GRAD = 2 /* v *= gelu'(aux[m,n])     (dgrad fused with the erf-GELU backward) */
} vb_aux_mode;

typedef struct vb_gemm_args {
  const void* a;       /* bf16 */
  const void* b;       /* bf16 */
  void* d;             /* bf16 [M,N] (d_is_f32=0) or fp32 [M,N] (d_is_f32=1) */
  void* d_preact;      /* optional bf16 [M,N]: value before aux/activation (GELU input kept for backward) */
  const float* scale;  /* optional fp32 [N]: v = acc*scale[n]   (folded BatchNorm) */
  const float* bias;   /* optional fp32 [N]: v += bias[n] */
  const void* aux;     /* optional bf16 [M,N], see aux_mode */
  int64_t lda, ldb, ldd, ld_preact, ld_aux; /* elements */
  int32_t m, n, k;
  int32_t a_mn_major, b_mn_major;
  int32_t d_is_f32;
  int32_t accumulate;  /* fp32 output only: D += result (gradient accumulation / split-K) */
  int32_t act;         /* vb_act */
  int32_t aux_mode;    /* vb_aux_mode */
  int32_t block_n;     /* 0 = auto; else 64 | 128 | 256 */
  int32_t splits;      /* 0 = auto (1 unless fp32+accumulate); >1 requires d_is_f32 && accumulate */
  int32_t max_ctas;    /* 0 = all SMs; otherwise cap the persistent grid (stream co-scheduling) */
} vb_gemm_args;

int vb_gemm_bf16(const vb_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VILBERT_B200_H */
