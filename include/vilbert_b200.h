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
  VB_AUX_MUL_GELU_GRAD = 2,/* v *= gelu'(aux[m,n])     (dgrad fused with the erf-GELU backward) */
  VB_AUX_MUL = 3           /* v *= aux[m,n]            (the same dgrad when the forward stored gelu' itself: preact_grad) */
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
  int32_t block_n;     /* 0 = auto; else a preferred tile width, a multiple of 32 in [64, 256] (ignored when illegal for the layout) */
  int32_t splits;      /* 0 = auto (1 unless fp32+accumulate); >1 requires d_is_f32 && accumulate */
  int32_t max_ctas;    /* 0 = all SMs; otherwise cap the persistent grid (stream co-scheduling) */
  int32_t b_streamed;  /* B is read once per step (a weight matrix): load it with the L2 evict-first hint */
  int32_t d_streamed;  /* D is not re-read soon (a weight gradient): store it with the L2 evict-first hint */
  /* Implicit-GEMM convolution (conv_kh > 0; nn.Conv2d of the ResNet trunk, resnet152_roi.py:49-74): `a` is the bf16 NHWC
   * activation [conv_n, conv_h, conv_w, conv_c] (contiguous, conv_c % 64 == 0), `b` the weight [n, conv_kh*conv_kw*conv_c] with
   * column (ky*kw + kx)*c + ci, d the NHWC output [conv_n*ho*wo, n]; m = conv_n*ho*wo, k = conv_kh*conv_kw*conv_c, lda unused.
   * The A tiles are fetched by TMA in im2col mode (zero padding by out-of-bounds fill): no [pixels, kh*kw*c] buffer exists. */
  int32_t conv_n, conv_h, conv_w, conv_c, conv_kh, conv_kw, conv_stride, conv_pad;
  /* With act = VB_ACT_GELU and d_preact: store GELU'(pre-activation) in d_preact instead of the pre-activation.  The forward
   * epilogue has the tanh of the GELU anyway, so the derivative costs a few FMAs there and the backward GEMM (VB_AUX_MUL) only
   * multiplies (vilbert_facebook_arch.py:184-185 and autograd thereof). */
  int32_t preact_grad;
} vb_gemm_args;

int vb_gemm_bf16(const vb_gemm_args* args, void* stream);
/* Profiling aid (tools/gemm_trace.py): when non-NULL, every CTA of later vb_gemm_bf16 launches writes 24 int64 clock64()
 * stamps (entry, prologue done, dependency wait done, first load, loads done, first operands landed, MMAs issued, first /
 * last accumulator ready, stores issued, staging drained, exit) to device_buffer[cta*24 ..].  NULL switches it off. */
int vb_gemm_set_trace(void* device_buffer);
/* Experiment knobs of the tile picker / kernel (the VB_GEMM_* environment variables, settable at run time by the profiling
 * tools): "occ1", "max_bn", "np", "cg", "debug_mode", "stages", "no_l2_hints".  Not for production use. */
int vb_gemm_set_knob(const char* name, int value);
/* Diagnostic (tools/gemm_occupancy.py): resident blocks per SM and co-resident 2-CTA clusters the runtime reports for the
 * narrow-tile pair kernel at `smem_bytes` of dynamic shared memory. */
int vb_gemm_debug_occupancy(int smem_bytes, int* blocks_per_sm, int* clusters);

/* ------------------------------------------------------------------------------------------------
 * Gradient exchange over NVLink / NVSwitch (data-parallel training: the one exchange step of the path; the
 * reference itself is single-GPU).  In-place mean all-reduce of the bf16 range [lo, hi) (elements, multiples of 8) of a
 * symmetric buffer: three launches on `stream` (rank barrier, slice reduce + broadcast, rank barrier).
 *   mc_base     multicast mapping of the buffer (NVLS: multimem.ld_reduce / multimem.st), or NULL
 *   peer_bases  DEVICE array of `world` pointers to every rank's mapping of the buffer (used when mc_base is NULL)
 *   flag_ptrs   DEVICE array of `world` pointers to every rank's flag pad (uint32, zero-initialised, flag_slots * world words)
 *   ctas        CTAs of the reduce kernel (0 = default)
 *   out_mc_base / out_peer_bases   optional SECOND symmetric buffer of fp32: when given, the mean of bf16 elements [lo, hi) is
 *               broadcast as fp32 into elements [out_lo, out_lo + hi - lo) of it on every rank (the bf16 buffer is left as it
 *               was); when both are NULL the result replaces the bf16 range in place.
 * vb_rank_barrier is the hand-shake alone (slot in [0, flag_slots)). */
typedef struct vb_exchange_args {
  void* mc_base;
  const void* peer_bases;
  const void* flag_ptrs;
  int32_t rank, world;
  int32_t flag_slots;
  int32_t ctas;
  void* out_mc_base;
  const void* out_peer_bases;
} vb_exchange_args;
int vb_allreduce_mean_bf16(const vb_exchange_args* args, int64_t lo, int64_t hi, int64_t out_lo, void* stream);
int vb_rank_barrier(const vb_exchange_args* args, int32_t slot, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-wise bandwidth kernels (one warp per row, 16-byte accesses, fp32 math on bf16 storage).
 * Dropout masks are a pure function of (*seed, site, element index): forward and backward regenerate
 * them, nothing is stored.  `seed` is a DEVICE pointer so that a captured CUDA graph sees a new seed on
 * every replay.
 * ---------------------------------------------------------------------------------------------- */
typedef enum vb_dtype { VB_DT_F32 = 0, VB_DT_I32 = 1, VB_DT_I64 = 2, VB_DT_BF16 = 3 } vb_dtype;

/* y = LayerNorm(dropout_in(x) + res) * gamma + beta, then dropout_out  — BertLayerNorm (eps inside the sqrt, biased
 * variance) fused with the dropout and residual add that precede it:
 *   models/vilbert_facebook_arch.py:63-76 (BertLayerNorm), :156-160 (BertSelfOutput), :197-201 (BertOutput),
 *   :329-336 (BiOutput), :100-104 (VisualEmbeddings: x = image term, res = location term, dropout_out).
 * Backward: dx = grad wrt x (masked), dres = grad wrt res, dgamma/dbeta/dbias (= column sum of dx, the bias gradient of
 * the dense layer that produced x) are ATOMICALLY ACCUMULATED into fp32 — zero them first. */
typedef struct vb_layernorm_args {
  const void* x;        /* bf16 [m,h] */
  const void* res;      /* bf16 [m,h] or NULL */
  const float* gamma;   /* fp32 [h] */
  const float* beta;    /* fp32 [h] */
  void* y;              /* bf16 [m,h]            (forward) */
  float* mean;          /* fp32 [m]  written by forward, read by backward */
  float* rstd;          /* fp32 [m] */
  const void* dy;       /* bf16 [m,h]            (backward) */
  void* dx;             /* bf16 [m,h] or NULL */
  void* dres;           /* bf16 [m,h] or NULL */
  float* dgamma;        /* fp32 [h] or NULL */
  float* dbeta;         /* fp32 [h] or NULL */
  float* dbias;         /* fp32 [h] or NULL */
  int64_t ldx, ldres, ldy, lddy, lddx, lddres; /* elements */
  int32_t m, h;         /* h in {256,512,768,1024,2048} */
  float eps;
  float p_in, p_out;    /* dropout probabilities (0 = off) */
  uint32_t site_in, site_out;
  const uint64_t* seed; /* device pointer, may be NULL when both p are 0 */
  /* fp32 residual stream (optional): the residual is read from res_f32 instead of res when given, and the output is ALSO
   * written in fp32 to y_f32, so that consecutive blocks hand the residual over unrounded (what PyTorch's bf16 autocast does:
   * LayerNorm and the residual add stay fp32, only the GEMM operands are bf16).  Leading dimensions in elements. */
  const float* res_f32; float* y_f32; int64_t ldres_f32, ldy_f32;
} vb_layernorm_args;
int vb_layernorm_fwd(const vb_layernorm_args* args, void* stream);
int vb_layernorm_bwd(const vb_layernorm_args* args, void* stream);

/* Text embeddings: y = dropout(LayerNorm(word[ids] + type[type_ids] + pos[0..t)))  — transformers BertEmbeddings.forward,
 * called at models/vilbert_facebook_arch.py:524.  Tables are the fp32 master parameters.  Backward scatter-adds into the
 * fp32 table gradients (atomics; word row 0 = padding_idx receives nothing). */
typedef struct vb_embed_args {
  const int32_t* ids;       /* [b*t] */
  const int32_t* type_ids;  /* [b*t] or NULL (all zero) */
  const float* word;        /* [vocab,h] */
  const float* pos;         /* [>=t,h] */
  const float* type;        /* [2,h] */
  const float* gamma;
  const float* beta;
  void* y;                  /* bf16 [b*t,h] */
  float* mean;
  float* rstd;
  const void* dy;           /* bf16 [b*t,h]  (backward) */
  float* dword; float* dpos; float* dtype; float* dgamma; float* dbeta;  /* each may be NULL (frozen) */
  int32_t b, t, h, vocab;
  float eps, p_out;
  uint32_t site_out;
  const uint64_t* seed;
  float* y_f32;             /* optional fp32 copy of y [b*t,h] (residual stream) */
} vb_embed_args;
int vb_embed_text_fwd(const vb_embed_args* args, void* stream);
int vb_embed_text_bwd(const vb_embed_args* args, void* stream);

/* out[n] += sum_m x[m,n]  (bias gradients; autograd of nn.Linear bias) */
int vb_colsum_bf16(const void* x, int64_t ld, int32_t m, int32_t n, float* out, void* stream);
/* fp32 -> bf16 (inputs; weight shadows).  The multi form converts many parameter tensors in one launch: `segs` is a
 * device array of {const float* src; bf16* dst; int64 n}, block i converts 8192 elements of segs[block_seg[i]] starting
 * at block_off[i]. */
int vb_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
/* bf16 -> fp32 (data-parallel: gradient buckets are all-reduced in bf16 and widened back into the fp32 gradient buffer) */
int vb_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream);
int vb_cast_f32_bf16_multi(const void* segs, const int32_t* block_seg, const int64_t* block_off, int32_t num_blocks,
                           void* stream);
/* out = (1.0f - (float)mask) * -10000.0f, bit-exact with models/vilbert_facebook_arch.py:530-540 */
int vb_mask_bias(const void* mask, int32_t mask_dtype, float* out, int32_t n, void* stream);
/* int64 ids / labels -> int32 with a range check ([lo,hi); *err_flag = 1 on violation, like the index error torch raises) */
int vb_i64_to_i32(const int64_t* src, int32_t* dst, int32_t n, int32_t lo, int32_t hi, int32_t* err_flag, void* stream);
/* y = dropout(x) element-wise (classifier nn.Dropout(0.1), models/vilbert_facebook_arch.py:573,576) */
int vb_dropout_bf16(const void* x, void* y, int64_t n, float p, uint32_t site, const uint64_t* seed, void* stream);
/* *seed = lcg(*seed): one tiny launch at the head of every training forward (dropout mask stream) */
int vb_seed_advance(uint64_t* seed, void* stream);
/* the same, and the new value is also written to *snapshot: every batch geometry ("plan") keeps the seed its own last
 * forward drew, so that its backward regenerates the right masks whatever other forwards ran in between */
int vb_seed_advance_to(uint64_t* seed, uint64_t* snapshot, void* stream);

/* Batch staging in ONE launch: what ViLBERTForClassification.forward (models/vilbert_facebook_arch.py:610-641) receives ->
 * the static buffers the captured graphs read.  Segment kinds:
 *   VB_STAGE_INDEX     int64 | int32 -> int32 with the range check [lo, hi) that nn.Embedding / nn.CrossEntropyLoss apply;
 *                      a violation ORs `err_bit` into *err_flag (the host raises) and stores `lo` (nothing downstream can
 *                      index out of bounds); with err_bit = VB_STAGE_ERR_LABEL the value -100 (CrossEntropyLoss's
 *                      ignore_index) is legal and kept
 *   VB_STAGE_MASK      int64 | int32 | fp32 attention mask -> (1.0f - m) * -10000.0f, bit-exact with :530-540
 *   VB_STAGE_FEAT      fp32 | bf16 region features -> bf16 (round to nearest even)
 *   VB_STAGE_COPY_F32  fp32 -> fp32 (the 5-d boxes) */
#define VB_STAGE_MAX_SEGS 8
#define VB_IGNORE_INDEX (-100)
typedef enum vb_stage_kind { VB_STAGE_INDEX = 0, VB_STAGE_MASK = 1, VB_STAGE_FEAT = 2, VB_STAGE_COPY_F32 = 3 } vb_stage_kind;
typedef enum vb_stage_err { VB_STAGE_ERR_ID = 1, VB_STAGE_ERR_TYPE = 2, VB_STAGE_ERR_LABEL = 4 } vb_stage_err;
typedef struct vb_stage_seg {
  const void* src;
  void* dst;
  int64_t n;        /* elements; 0 = segment absent */
  int32_t kind;     /* vb_stage_kind */
  int32_t dtype;    /* vb_dtype of src */
  int32_t lo, hi;   /* VB_STAGE_INDEX: legal range [lo, hi) */
  int32_t err_bit;  /* VB_STAGE_INDEX: vb_stage_err bit */
} vb_stage_seg;
int vb_stage_batch(const vb_stage_seg* segs, int32_t nseg, int32_t* err_flag, void* stream);
/* dx = dy * act'(y) for tanh (BertPooler :407) and ReLU (classifier :575), through the activation output y */
int vb_act_bwd_bf16(const void* dy, const void* y, void* dx, int64_t n, int32_t act, void* stream);
/* image_location_embeddings (Linear(5,1024), models/vilbert_facebook_arch.py:92-94,102): forward term and its gradients */
int vb_loc_embed_fwd(const float* loc, const float* w, const float* b, void* out, int32_t m, int32_t n, int32_t kdim,
                     void* stream);
int vb_loc_embed_bwd(const void* ds, const float* loc, float* dw, float* db, int32_t m, int32_t n, int32_t kdim,
                     void* stream);
/* classifier tail Linear(1024,num_labels) + CrossEntropyLoss(mean), models/vilbert_facebook_arch.py:577, 637-639.
 * fwd: logits fp32 [b,c], probs = softmax(logits), *loss (labels may be NULL -> loss 0).
 * bwd: dlogits = *dloss * (probs - onehot)/nv + dlogits_ext;  dw[c,k], db[c] (overwritten), dh bf16 [b,k].
 * Labels equal to VB_IGNORE_INDEX (-100, the default ignore_index of nn.CrossEntropyLoss) contribute neither loss nor
 * gradient and nv counts the others (all ignored -> loss NaN, as torch). */
int vb_cls_ce_fwd(const void* h, const float* w, const float* bias, const int32_t* labels, float* logits, float* probs,
                  float* loss, int32_t bsz, int32_t kdim, int32_t c, void* stream);
int vb_cls_ce_bwd(const void* h, const float* w, const int32_t* labels, const float* probs, const float* dloss,
                  const float* dlogits_ext, float* dw, float* db, void* dh, int32_t bsz, int32_t kdim, int32_t c,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused softmax attention on tcgen05/TMEM: one CTA per (sample, head), sq, sk <= 128 per call, d in {64,128}.
 *   out = dropout(softmax(q k^T * scale + mask_bias[b, key])) v
 * Serves BertSelfAttention (models/vilbert_facebook_arch.py:126-144) and both directions of BiAttention (:253-294):
 * q, k, v are independent strided views (pointer to the first head's column, row stride in elements, sq / sk rows
 * per sample).  lse [batch, heads, 128] fp32 is written by forward and read by backward.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vb_attn_args {
  const void* q; const void* k; const void* v;   /* bf16 */
  int64_t ldq, ldk, ldv;
  void* out; int64_t ldo;                        /* bf16 [batch*sq, heads*d] view */
  float* lse;
  const float* mask_bias;                        /* fp32 [batch, sk] or NULL */
  int32_t batch, heads, sq, sk, d;
  float scale;
  float p_drop; uint32_t site; const uint64_t* seed;
  const void* dout; int64_t lddo;                /* backward */
  void* dq; void* dk; void* dv;
  int64_t lddq, lddk, lddv;
  /* Block views, for sequences above 128 (0 = the block is the whole sequence): q / out / dout / dq address sample b at row
   * b * q_batch_rows, k / v / dk / dv at row b * k_batch_rows, mask_bias at b * bias_ld; rows past sq / sk are not touched. */
  int64_t q_batch_rows, k_batch_rows, bias_ld;
  const float* delta;                            /* backward of a key block: sum over ALL keys of P dP per row, fp32
                                                    [batch, heads, 128] (vb_attn_delta); NULL = this block holds all keys */
} vb_attn_args;
int vb_attention_fwd(const vb_attn_args* args, void* stream);
int vb_attention_bwd(const vb_attn_args* args, void* stream);

/* Longer sequences (e.g. 257 DINOv2 patch tokens as regions, BASELINE config 4) run as blocks of <= 128 queries x <= 128 keys
 * through the two entry points above; these three kernels join the blocks (flash-attention algebra, host loop in ops.py):
 *   vb_attn_merge : out = sum_j exp(lse_j - LSE) o_j, LSE = log sum_j exp(lse_j) over n_parts <= 4 key blocks; o_parts /
 *                   lse_parts are HOST arrays of device pointers (bf16 [.., heads*d] views with row stride ldp, fp32
 *                   [batch, heads, 128]); writes out (bf16, row stride ldo) and lse_out.
 *   vb_attn_delta : delta[b, h, row] = sum_d dout * out over one head's columns (= sum_k P dP over all keys).
 *   vb_sum_rows_bf16 : dst[r, c] = sum_p parts[p][r, c] (fp32 accumulation) for the dq / dk / dv partials of the blocks.
 * Row r of sample b lives at row b * batch_rows + r of every view; sq rows per sample are processed. */
int vb_attn_merge(const void* const* o_parts, const float* const* lse_parts, int32_t n_parts, int64_t ldp, void* out,
                  int64_t ldo, float* lse_out, int32_t batch, int32_t heads, int32_t sq, int64_t batch_rows, int32_t d,
                  void* stream);
int vb_attn_delta(const void* out, int64_t ldo, const void* dout, int64_t lddo, float* delta, int32_t batch, int32_t heads,
                  int32_t sq, int64_t batch_rows, int32_t d, void* stream);
int vb_sum_rows_bf16(const void* const* parts, int32_t n_parts, int64_t ldp, void* dst, int64_t ldd, int64_t rows,
                     int32_t width, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ResNet-152 RoI feature stage (models/feature_extractors/resnet152_roi.py).  Activations are NHWC bf16; every
 * convolution of torchvision's resnet152 (conv1..layer3 at :49-57 = forward_base, layer4 at :69-74 = forward_top) runs as
 * vb_gemm_bf16 over [pixels, kh*kw*Cin] with the eval-mode BatchNorm folded into `scale` / `bias`, the bottleneck's
 * identity as `aux` (VB_AUX_ADD) and VB_ACT_RELU.  The kernels below produce the GEMM operands and pool.
 * ---------------------------------------------------------------------------------------------- */
/* conv1 7x7/2 operand: fp32 NCHW image [n,3,h,w] -> bf16 [n*ho*wo, kpad], column (ky*kw+kx)*3+c, zero tail up to kpad */
int vb_stem_im2col(const float* img, void* y, int32_t n, int32_t h, int32_t w, int32_t kh, int32_t kw, int32_t stride,
                   int32_t pad, int32_t kpad, void* stream);
/* bf16 NHWC [n,h,w,c] -> bf16 [n*ho*wo, kh*kw*c], column (ky*kw+kx)*c+ci (3x3 convolutions; 1x1 stride-2 downsample) */
int vb_im2col_nhwc(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t kh, int32_t kw,
                   int32_t stride, int32_t pad, void* stream);
/* nn.MaxPool2d(k, stride, pad) on NHWC bf16 (resnet.maxpool, :52) */
int vb_maxpool_nhwc(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, int32_t stride,
                    int32_t pad, void* stream);
/* torchvision.ops.RoIPool((ph,pw), spatial_scale) (:126, :167-170): rois fp32 [r,5] = (batch index, x1, y1, x2, y2);
 * out bf16 [r,ph,pw,c]; argmax (optional) int32 [r,ph,pw,c] = iy*w+ix of the selected element, -1 for an empty bin.
 * Index arithmetic is bit-exact with torchvision (round-half-away, floor/ceil bins, clipping, empty bin -> 0). */
int vb_roi_pool_nhwc(const void* x, const float* rois, void* y, int32_t* argmax, int32_t num_rois, int32_t n, int32_t h,
                     int32_t w, int32_t c, int32_t ph, int32_t pw, float spatial_scale, void* stream);
/* torchvision.ops.roi_align(output_size, spatial_scale, sampling_ratio, aligned) as used by MultiScaleRoIAlign in
 * models/feature_extractors/fasterrcnn_resnet152.py:130-134 (7x7, sampling_ratio 2) */
int vb_roi_align_nhwc(const void* x, const float* rois, void* y, int32_t num_rois, int32_t n, int32_t h, int32_t w,
                      int32_t c, int32_t ph, int32_t pw, float spatial_scale, int32_t sampling_ratio, int32_t aligned,
                      void* stream);
/* AdaptiveAvgPool2d((1,1)) + flatten (:71-73): bf16 [r, s, c] -> fp32 [r, c] */
int vb_avgpool_nhwc(const void* x, float* out, int32_t r, int32_t s, int32_t c, void* stream);

/* Proposal selection (:251-293).  scores[i] = 1 - |((x2-x1)/img_w) * ((y2-y1)/img_h) - target_area| in the reference's fp32
 * operation order (:262-270); vb_nms = torchvision.ops.nms(boxes, scores, iou_threshold) (:277) with the CPU kernel's
 * semantics: stable descending sort of the scores, greedy suppression of IoU > threshold (fp32 IoU compared with the DOUBLE
 * threshold, as torchvision does), kept indices in score order.
 * Bit-exact: the scores of translated copies of one box size tie exactly and the tie order decides which boxes survive.
 * boxes fp32 [n,4] (x1,y1,x2,y2); workspace int32 [2n]; keep int32 [n]; *num_keep int32 (all device). */
int vb_box_area_score(const float* boxes, int32_t n, float img_w, float img_h, float target_area, float* scores, void* stream);
int vb_nms(const float* boxes, const float* scores, int32_t n, double iou_threshold, int32_t* workspace, int32_t* keep,
           int32_t* num_keep, void* stream);

/* Visual Genome Faster R-CNN extractor (next-row f-4; models/feature_extractors/fasterrcnn_vg.py).
 *   vb_rowmax_f32     : out[r] = max over columns [col_begin, col_end) of fp32 x[rows, ld] -- the proposal score
 *                       ``cls_scores[:, 1:].max(dim=1)[0]`` of _score_proposals (:345-365; column 0 = background).
 *   vb_select_regions : _select_top_regions + _pad_regions + _extract_roi_features' row choice + _normalize_boxes
 *                       (:367-411, 436-469) on the device, no host round trip: output region r is candidate
 *                       keep[min(r, *num_keep - 1)] (vb_nms leaves the survivors in descending score order, so the
 *                       reference's top-k over them is their first `regions` entries; among exactly tied scores torch.topk's
 *                       order is unspecified and the stable order is kept).  Writes, each optional (NULL): boxes fp32
 *                       [regions,4], spatial fp32 [regions,5] = (x1/img_w, y1/img_h, x2/img_w, y2/img_h clamped to [0,1],
 *                       area) bit-exact with the reference's fp32 tensor ops, index int32 [regions], and feat_dst fp32
 *                       [regions, feat_dim] = rows of feat_src [n, feat_dim], rois fp32 [regions,5] = (batch_index, box): the
 *                       vb_roi_pool_nhwc operand of the chosen boxes.  *num_keep == 0 leaves the outputs untouched.  box_div:
 *                       the boxes are divided by it (fp32, one rounding) before they are normalised -- the resize factor of
 *                       fasterrcnn_vg_rpn.py:432 (``boxes / scale``); 1.0 is exact. */
int vb_rowmax_f32(const float* x, int32_t rows, int32_t ld, int32_t col_begin, int32_t col_end, float* out, void* stream);
int vb_select_regions(const float* candidates, const int32_t* keep, const int32_t* num_keep, int32_t regions, float img_w,
                      float img_h, const float* feat_src, int32_t feat_dim, float* boxes, float* spatial, float* feat_dst,
                      int32_t* index, float* rois, float batch_index, float box_div, void* stream);

/* RPN-proposal variant of the Visual Genome extractor (models/feature_extractors/fasterrcnn_vg_rpn.py), post-processing on the
 * device with no host read between the steps (counts stay in device memory):
 *   vb_rpn_decode     : RPN.forward after the convolutions (:78-104) + _generate_anchors / _apply_deltas (:106-174) +
 *                       clip_boxes_to_image + the min-size test of _filter_proposals (:444-450).  heads fp32 [fh*fw, ld]: columns
 *                       [0, 2A) objectness logits (anchor k: 2k = background, 2k+1 = foreground), [2A, 6A) box deltas (4 per
 *                       anchor); base_anchors = HOST array [A,4] (A <= 16).  boxes fp32 [fh*fw*A, 4] (anchor fastest), scores
 *                       = softmax foreground probability, -inf where the clipped box is below min_size; *num_valid = the rest.
 *   vb_rank_sort_desc : order[rank] = i, stable descending (torch.sort(descending=True, stable=True)); elements at or beyond
 *                       *limit (optional device int) count as -inf.  The reference's torch.topk (:428, :455-458) is its prefix;
 *                       among exactly tied scores topk's order is unspecified and the stable one is kept.
 *   vb_gather_sorted  : the first min(cap, *num_valid) boxes / scores in that order (rows beyond: zero box, -inf); *count.
 *   vb_nms_sorted     : torchvision.ops.nms over boxes ALREADY in descending score order (:461), stopping at max_keep survivors
 *                       (:464-465); keep int32 [max_keep] = positions, *num_keep.  *count <= 8192. */
int vb_rpn_decode(const float* heads, int32_t ld, int32_t fh, int32_t fw, int32_t num_anchors, const float* base_anchors,
                  float stride, float img_h, float img_w, float min_size, float* boxes, float* scores, int32_t* num_valid,
                  void* stream);
int vb_rank_sort_desc(const float* scores, int32_t n, const int32_t* limit, int32_t* order, void* stream);
int vb_gather_sorted(const float* boxes, const float* scores, const int32_t* order, const int32_t* num_valid, int32_t cap,
                     float* out_boxes, float* out_scores, int32_t* count, void* stream);
int vb_nms_sorted(const float* boxes, const int32_t* count, double iou_threshold, int32_t max_keep, int32_t* keep,
                  int32_t* num_keep, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DINOv2 multi-layer fusion tail (next-row f-2; models/feature_extractors/dinov2_multilayer.py:342-381).
 *   vb_bilinear_concat : torch.cat(layer_features, -1) -> [g x g grid] -> F.interpolate(size=(t,t), bilinear,
 *                        align_corners=False) -> [batch*t*t, L*h] bf16.  layers: HOST array of L device pointers to fp32
 *                        [batch, tokens, h] (strides in elements; first_token = 1 skips the CLS token, :320).
 *   vb_gelu_bf16       : nn.GELU() (erf) between the LayerNorm and the second Linear of `projection` (:250-255).
 * The two Linears are vb_gemm_bf16, the LayerNorm is vb_layernorm_fwd (eps 1e-5).
 * ---------------------------------------------------------------------------------------------- */
int vb_bilinear_concat(const float* const* layers, int32_t num_layers, void* out, int32_t batch, int32_t grid, int32_t target,
                       int32_t h, int64_t batch_stride, int64_t token_stride, int32_t first_token, void* stream);
int vb_gelu_bf16(const void* x, void* y, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Feature-store ingest (next-row f-3; pipelines/data_processing/lmdb_dataset.py:143-208).
 *   vb_lmdb_regions : one batch of detectron.lmdb records, already in HBM as raw fp32 arrays, becomes what the encoder
 *                     consumes.  features fp32 [n_features] (= rows * feature_dim, multiple of 8) -> features_bf16
 *                     (round-to-nearest-even, the A operand of `image_embeddings`); boxes fp32 [rows, box_stride >= 4]
 *                     (x1, y1, x2, y2, ...) -> spatial fp32 [rows, 5] = [x1/box_div, y1/box_div, x2/box_div, y2/box_div,
 *                     ((x2-x1)*(y2-y1))/area_div], the float32 operation order of `_process_boxes` (:189-208; box_div
 *                     1000, area_div 1e6), bit-exact.  rows = 0 skips the boxes (HDF5-layout stores keep final spatial
 *                     rows, precomputed_dataset.py:90-92); n_features = 0 skips the cast.  One launch.
 * ---------------------------------------------------------------------------------------------- */
int vb_lmdb_regions(const float* features, void* features_bf16, int64_t n_features, const float* boxes, float* spatial,
                    int32_t rows, int32_t box_stride, float box_div, float area_div, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused optimizer step over the flat parameter / gradient buffers (next-row f-1 of SURVEY.md §8): replaces
 * torch.nn.utils.clip_grad_norm_(params, max_norm) + torch.optim.AdamW.step() of pipelines/model_training/nodes.py:795-799.
 *   vb_grad_sumsq : *acc += sum(grad[i]^2)  (fp64 device accumulator, zero it first; call once per contiguous range)
 *   vb_adamw_step : g = grad * min(max_norm / (sqrt(*grad_sumsq) + 1e-6), 1) when max_norm > 0; decoupled weight decay; Adam
 *                   moments; bias-corrected update (torch/optim/adam.py::_single_tensor_adam, fp32); and the bf16 shadow of
 *                   the first shadow_n elements (the GEMM weights) rewritten in the same pass.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vb_adamw_args {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;   /* fp32 [n] */
  void* shadow;                 /* bf16 [shadow_n] or NULL */
  int64_t n, shadow_n;
  const double* grad_sumsq;     /* device scalar from vb_grad_sumsq; required when max_norm > 0 */
  float max_norm;               /* <= 0: no clipping */
  float lr, beta1, beta2, eps, weight_decay;
  int32_t step;                 /* 1-based step count (bias corrections) */
} vb_adamw_args;
int vb_grad_sumsq(const float* grad, int64_t n, double* acc, void* stream);
int vb_adamw_step(const vb_adamw_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VILBERT_B200_H */
