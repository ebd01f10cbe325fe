"""ctypes binding of libvilbert_b200.so (the C ABI declared in include/vilbert_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
fallback: if it is missing, or a call returns a non-zero status, we raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VB_LIB: an alternative build of the SAME library (e.g. the -DVB_GEMM_TRACE instrumentation build used by tools/gemm_trace.py)
LIB_PATH = os.environ.get("VB_LIB") or os.path.join(_HERE, "libvilbert_b200.so")

ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH = 0, 1, 2, 3
AUX_NONE, AUX_ADD, AUX_MUL_GELU_GRAD, AUX_MUL = 0, 1, 2, 3


class VbError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("d", C.c_void_p), ("d_preact", C.c_void_p),
        ("scale", C.c_void_p), ("bias", C.c_void_p), ("aux", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64), ("ld_preact", C.c_int64), ("ld_aux", C.c_int64),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("d_is_f32", C.c_int32), ("accumulate", C.c_int32), ("act", C.c_int32), ("aux_mode", C.c_int32),
        ("block_n", C.c_int32), ("splits", C.c_int32), ("max_ctas", C.c_int32),
        ("b_streamed", C.c_int32), ("d_streamed", C.c_int32),
        ("conv_n", C.c_int32), ("conv_h", C.c_int32), ("conv_w", C.c_int32), ("conv_c", C.c_int32),
        ("conv_kh", C.c_int32), ("conv_kw", C.c_int32), ("conv_stride", C.c_int32), ("conv_pad", C.c_int32),
        ("preact_grad", C.c_int32),
    ]


class LayerNormArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("res", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("y", C.c_void_p),
        ("mean", C.c_void_p), ("rstd", C.c_void_p), ("dy", C.c_void_p), ("dx", C.c_void_p), ("dres", C.c_void_p),
        ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("dbias", C.c_void_p),
        ("ldx", C.c_int64), ("ldres", C.c_int64), ("ldy", C.c_int64), ("lddy", C.c_int64), ("lddx", C.c_int64),
        ("lddres", C.c_int64),
        ("m", C.c_int32), ("h", C.c_int32), ("eps", C.c_float), ("p_in", C.c_float), ("p_out", C.c_float),
        ("site_in", C.c_uint32), ("site_out", C.c_uint32), ("seed", C.c_void_p),
        ("res_f32", C.c_void_p), ("y_f32", C.c_void_p), ("ldres_f32", C.c_int64), ("ldy_f32", C.c_int64),
    ]


class EmbedArgs(C.Structure):
    _fields_ = [
        ("ids", C.c_void_p), ("type_ids", C.c_void_p), ("word", C.c_void_p), ("pos", C.c_void_p), ("type", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("y", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("dy", C.c_void_p), ("dword", C.c_void_p), ("dpos", C.c_void_p), ("dtype", C.c_void_p), ("dgamma", C.c_void_p),
        ("dbeta", C.c_void_p),
        ("b", C.c_int32), ("t", C.c_int32), ("h", C.c_int32), ("vocab", C.c_int32), ("eps", C.c_float),
        ("p_out", C.c_float), ("site_out", C.c_uint32), ("seed", C.c_void_p), ("y_f32", C.c_void_p),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("lse", C.c_void_p), ("mask_bias", C.c_void_p),
        ("batch", C.c_int32), ("heads", C.c_int32), ("sq", C.c_int32), ("sk", C.c_int32), ("d", C.c_int32),
        ("scale", C.c_float), ("p_drop", C.c_float), ("site", C.c_uint32), ("seed", C.c_void_p),
        ("dout", C.c_void_p), ("lddo", C.c_int64), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("lddq", C.c_int64), ("lddk", C.c_int64), ("lddv", C.c_int64),
        ("q_batch_rows", C.c_int64), ("k_batch_rows", C.c_int64), ("bias_ld", C.c_int64), ("delta", C.c_void_p),
    ]


class AdamWArgs(C.Structure):
    _fields_ = [
        ("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("shadow", C.c_void_p),
        ("n", C.c_int64), ("shadow_n", C.c_int64), ("grad_sumsq", C.c_void_p), ("max_norm", C.c_float),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
        ("step", C.c_int32),
    ]


class StageSeg(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n", C.c_int64), ("kind", C.c_int32), ("dtype", C.c_int32),
                ("lo", C.c_int32), ("hi", C.c_int32), ("err_bit", C.c_int32)]


class ExchangeArgs(C.Structure):
    _fields_ = [("mc_base", C.c_void_p), ("peer_bases", C.c_void_p), ("flag_ptrs", C.c_void_p), ("rank", C.c_int32),
                ("world", C.c_int32), ("flag_slots", C.c_int32), ("ctas", C.c_int32), ("out_mc_base", C.c_void_p),
                ("out_peer_bases", C.c_void_p)]


DT_F32, DT_I32, DT_I64, DT_BF16 = 0, 1, 2, 3
STAGE_INDEX, STAGE_MASK, STAGE_FEAT, STAGE_COPY_F32 = 0, 1, 2, 3
STAGE_ERR_ID, STAGE_ERR_TYPE, STAGE_ERR_LABEL = 1, 2, 4
IGNORE_INDEX = -100

_lib = None


def lib() -> C.CDLL:
    """Load the kernel library, failing loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VbError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU or PyTorch fallback.")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(l: C.CDLL) -> None:
    l.vb_abi_version.restype = C.c_int
    l.vb_last_error.restype = C.c_char_p
    l.vb_build_info.restype = C.c_char_p
    l.vb_gemm_bf16.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    l.vb_gemm_bf16.restype = C.c_int
    vp, i32, i64, f32, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32
    sigs = {
        "vb_layernorm_fwd": [C.POINTER(LayerNormArgs), vp], "vb_layernorm_bwd": [C.POINTER(LayerNormArgs), vp],
        "vb_embed_text_fwd": [C.POINTER(EmbedArgs), vp], "vb_embed_text_bwd": [C.POINTER(EmbedArgs), vp],
        "vb_attention_fwd": [C.POINTER(AttnArgs), vp], "vb_attention_bwd": [C.POINTER(AttnArgs), vp],
        "vb_gemm_set_trace": [vp],
        "vb_gemm_set_knob": [C.c_char_p, i32],
        "vb_colsum_bf16": [vp, i64, i32, i32, vp, vp],
        "vb_cast_f32_bf16": [vp, vp, i64, vp],
        "vb_cast_bf16_f32": [vp, vp, i64, vp],
        "vb_cast_f32_bf16_multi": [vp, vp, vp, i32, vp],
        "vb_mask_bias": [vp, i32, vp, i32, vp],
        "vb_i64_to_i32": [vp, vp, i32, i32, i32, vp, vp],
        "vb_dropout_bf16": [vp, vp, i64, f32, u32, vp, vp],
        "vb_seed_advance": [vp, vp],
        "vb_seed_advance_to": [vp, vp, vp],
        "vb_stage_batch": [C.POINTER(StageSeg), i32, vp, vp],
        "vb_allreduce_mean_bf16": [C.POINTER(ExchangeArgs), i64, i64, i64, vp],
        "vb_rank_barrier": [C.POINTER(ExchangeArgs), i32, vp],
        "vb_act_bwd_bf16": [vp, vp, vp, i64, i32, vp],
        "vb_loc_embed_fwd": [vp, vp, vp, vp, i32, i32, i32, vp],
        "vb_loc_embed_bwd": [vp, vp, vp, vp, i32, i32, i32, vp],
        "vb_cls_ce_fwd": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "vb_cls_ce_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "vb_stem_im2col": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
        "vb_im2col_nhwc": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
        "vb_maxpool_nhwc": [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
        "vb_roi_pool_nhwc": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, vp],
        "vb_roi_align_nhwc": [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, i32, i32, vp],
        "vb_avgpool_nhwc": [vp, vp, i32, i32, i32, vp],
        "vb_bilinear_concat": [vp, i32, vp, i32, i32, i32, i32, i64, i64, i32, vp],
        "vb_gelu_bf16": [vp, vp, i64, vp],
        "vb_grad_sumsq": [vp, i64, vp, vp],
        "vb_adamw_step": [C.POINTER(AdamWArgs), vp],
        "vb_box_area_score": [vp, i32, f32, f32, f32, vp, vp],
        "vb_nms": [vp, vp, i32, C.c_double, vp, vp, vp, vp],
        "vb_rowmax_f32": [vp, i32, i32, i32, i32, vp, vp],
        "vb_select_regions": [vp, vp, vp, i32, f32, f32, vp, i32, vp, vp, vp, vp, vp, f32, f32, vp],
        "vb_rpn_decode": [vp, i32, i32, i32, i32, vp, f32, f32, f32, f32, vp, vp, vp, vp],
        "vb_rank_sort_desc": [vp, i32, vp, vp, vp],
        "vb_gather_sorted": [vp, vp, vp, vp, i32, vp, vp, vp, vp],
        "vb_nms_sorted": [vp, vp, C.c_double, i32, vp, vp, vp],
        "vb_lmdb_regions": [vp, vp, i64, vp, vp, i32, i32, f32, f32, vp],
        "vb_attn_merge": [vp, vp, i32, i64, vp, i64, vp, i32, i32, i32, i64, i32, vp],
        "vb_attn_delta": [vp, i64, vp, i64, vp, i32, i32, i32, i64, i32, vp],
        "vb_sum_rows_bf16": [vp, i32, i64, vp, i64, i64, i32, vp],
    }
    for name, argtypes in sigs.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int


_launches = 0


def launch_count() -> int:
    """Number of successful kernel-launching ABI calls made from this process (each is exactly one kernel launch)."""
    return _launches


def check(rc: int, what: str) -> None:
    global _launches
    _launches += 1
    if rc != 0:
        msg = lib().vb_last_error().decode("utf-8", "replace")
        raise VbError(f"{what} failed with status {rc}: {msg}")


def exported_symbols_in_header() -> list[str]:
    """Names of every function declared in include/vilbert_b200.h (for the ABI test)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "vilbert_b200.h")
    text = open(hdr).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", text)))
