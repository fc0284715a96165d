"""ctypes binding of libhoisdf_b200.so (declared in include/hoisdf_b200.h).

There is NO fallback: if the CUDA library is missing or does not export a declared symbol, importing
this module raises.  Build it with `python -m hoisdf_b200.csrc.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhoisdf_b200.so")

ABI_VERSION = 39
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2      # ACT_SIGMOID: hoisdf_linear_narrow_split_fwd only
GATHER_CONCAT, GATHER_SUM = 0, 1

c_float_p = C.c_void_p  # device pointers are passed as integers (tensor.data_ptr())
i64, i32, f32, vp = C.c_int64, C.c_int32, C.c_float, C.c_void_p


class LinearArgs(C.Structure):
    _fields_ = [
        ("x", vp), ("ldx", i64), ("x_rows_per_batch", i64), ("x_batch_stride", i64),
        ("w", vp), ("ldw", i64), ("bias", vp), ("residual", vp),
        ("y", vp), ("ldy", i64), ("y_rows_per_batch", i64), ("y_batch_stride", i64),
        ("m", i64), ("n", i64), ("k", i64), ("act", i32), ("w_lo", vp), ("tf32_passes", i32),
    ]


class LinearH3Args(C.Structure):
    _fields_ = [
        ("x_hi", vp), ("x_lo", vp), ("ldx", i64), ("x_rows_per_batch", i64), ("x_batch_stride", i64),
        ("w_a", vp), ("w_b", vp), ("w_c", vp), ("ldw", i64), ("bias", vp), ("residual", vp),
        ("y", vp), ("ldy", i64), ("y_hi", vp), ("y_lo", vp), ("ldyh", i64),
        ("m", i64), ("n", i64), ("k", i64), ("act", i32), ("chunk_kb", i32),
        ("res_hi", vp), ("res_lo", vp), ("ldr", i64), ("single_pass", i32), ("w_scale", f32), ("y_scale", vp),
        ("split_k", i32),
    ]


class ConvH3Args(C.Structure):
    _fields_ = [
        ("x_hi", vp), ("x_lo", vp), ("batch", i64), ("in_h", i64), ("in_w", i64), ("cin", i64), ("ldx", i64),
        ("w_a", vp), ("w_b", vp), ("w_c", vp), ("ldw", i64), ("bias", vp),
        ("taps", i32), ("tap_dy", i32 * 16), ("tap_dx", i32 * 16), ("stride", i32),
        ("out_h", i64), ("out_w", i64), ("cout", i64),
        ("y", vp), ("y_hi", vp), ("y_lo", vp), ("y_sx", i64), ("y_sy", i64), ("y_sb", i64),
        ("act", i32), ("chunk_kb", i32),
        ("res_hi", vp), ("res_lo", vp), ("ldr", i64), ("single_pass", i32), ("w_scale", f32),
    ]


class Pyramid(C.Structure):
    _fields_ = [
        ("map", vp * 5), ("c", i32 * 5), ("h", i32 * 5), ("w", i32 * 5),
        ("levels", i32), ("img_h", i32), ("img_w", i32),
    ]


class SdfWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4",
                                   "w0_lo", "w1_lo", "w2_lo", "w3_lo")] + [("tf32_passes", i32)]


class SdfWeightsH3(C.Structure):
    _fields_ = [("w", (vp * 3) * 4), ("ldw", i64 * 4), ("b", vp * 4), ("w4", vp), ("b4", vp), ("chunk_kb", i32), ("single_pass", i32)]


class SdfChainArgs(C.Structure):
    _fields_ = [
        ("a0", vp), ("lda0", i64), ("x", vp), ("ldx", i64),
        ("lattice_index", vp), ("points", vp), ("bins", i32),
        ("w_s1", vp), ("ldw_s1", i64), ("b_s1", vp),
        ("w", vp * 4), ("ldw", i64 * 4), ("b", vp * 4), ("w4", vp), ("b4", vp),
        ("rows", i64), ("clamp", f32), ("out_sdf", vp),
        ("gmaps", vp), ("uv", vp), ("row_offsets", vp), ("batch", i64), ("rows_per_sample", i64), ("b_s0", vp),
    ]


class PyramidH(C.Structure):
    _fields_ = [("map", vp * 5), ("h", i32 * 5), ("w", i32 * 5), ("levels", i32), ("c", i32), ("img_h", i32), ("img_w", i32)]


class SdfInferArgs(C.Structure):
    _fields_ = [
        ("center", vp), ("cam_intr", vp), ("bbox", vp),
        ("sdf_scale", f32), ("bins", i32), ("batch", i64), ("num_points", i64), ("margin", i64), ("clamp", f32),
        ("gmaps", vp), ("gmaps16", vp), ("bias0", vp),
        ("s1_a", vp), ("s1_b", vp), ("s1_c", vp), ("ld_s1", i64), ("b_s1", vp), ("s1_scale", f32),
        ("dec", vp),
        ("workspace", vp), ("workspace_bytes", i64), ("max_rows", i64),
        ("planned", i32), ("chunk_counts", vp), ("offsets", vp), ("host_offsets", vp), ("n_f", vp),
        ("points", vp), ("sdf", vp), ("posenc", vp), ("sel_index", vp), ("status_flag", vp),
        ("screen_err", vp), ("screen_gap", vp), ("verified", vp),
        ("cand_sdf", vp), ("cand_index", vp), ("exact_sdf", vp), ("exact_index", vp), ("screen_rows", vp),
    ]


class H3Linear(C.Structure):
    _fields_ = [("a", vp), ("b", vp), ("c", vp), ("ld", i64), ("bias", vp), ("scale", f32)]


class EncoderLayer(C.Structure):
    _fields_ = [("qkv", H3Linear), ("out", H3Linear), ("lin1", H3Linear), ("lin2", H3Linear),
                ("norm1_g", vp), ("norm1_b", vp), ("norm2_g", vp), ("norm2_b", vp)]


class EncoderArgs(C.Structure):
    _fields_ = [("layers", C.POINTER(EncoderLayer)), ("num_layers", i32), ("heads", i32), ("d_ff", i64),
                ("inter_g", vp), ("inter_b", vp), ("batch", i64), ("seq", i64), ("x", vp),
                ("out", vp), ("out_hi", vp), ("out_lo", vp), ("ld_out", i64),
                ("inter", vp), ("inter_hi", vp), ("inter_lo", vp), ("ld_inter", i64),
                ("workspace", vp), ("workspace_bytes", i64)]


class DecoderLayer(C.Structure):
    _fields_ = [(n, H3Linear) for n in ("sa_qk", "sa_v", "sa_out", "ca_q", "ca_kv", "ca_out", "lin1", "lin2")] + \
               [(n, vp) for n in ("norm1_g", "norm1_b", "norm2_g", "norm2_b", "norm3_g", "norm3_b")]


class DecoderArgs(C.Structure):
    _fields_ = [("layers", C.POINTER(DecoderLayer)), ("num_layers", i32), ("heads", i32), ("d_ff", i64),
                ("norm_g", vp), ("norm_b", vp), ("batch", i64), ("queries", i64), ("seq", i64), ("kv_valid", i64),
                ("query_pos", vp), ("tgt_mask", vp), ("memory_hi", vp), ("memory_lo", vp), ("ld_memory", i64),
                ("hs", vp), ("workspace", vp), ("workspace_bytes", i64)]


class ManoModel(C.Structure):
    _fields_ = [(n, vp) for n in ("shapedirs", "posedirs", "v_template", "j_regressor", "weights", "hands_mean")]


# name -> (restype, argtypes); mirrors include/hoisdf_b200.h one to one
SIGNATURES = {
    "hoisdf_abi_version": (C.c_int, []),
    "hoisdf_status_string": (C.c_char_p, [C.c_int]),
    "hoisdf_linear_fwd": (C.c_int, [C.POINTER(LinearArgs), vp]),
    "hoisdf_split_tf32": (C.c_int, [vp, i64, vp, vp, vp]),
    "hoisdf_linear_h3_fwd": (C.c_int, [C.POINTER(LinearH3Args), vp]),
    "hoisdf_conv_h3_fwd": (C.c_int, [C.POINTER(ConvH3Args), vp]),
    "hoisdf_stem_im2col_split": (C.c_int, [vp, i64, i64, i64, vp, vp, i64, vp]),
    "hoisdf_maxpool3x3s2_split": (C.c_int, [vp, vp, i64, i64, i64, i64, i64, vp, vp, i64, vp]),
    "hoisdf_linear_narrow_split_fwd": (C.c_int, [vp, vp, i64, i64, vp, i64, vp, i64, i64, i32, vp, i64, vp]),
    "hoisdf_pack_h3": (C.c_int, [vp, i64, i64, i64, vp, vp, vp, i64, vp]),
    "hoisdf_split_rows": (C.c_int, [vp, i64, i64, i64, i64, vp, vp, i64, vp]),
    "hoisdf_join_rows": (C.c_int, [vp, vp, i64, i64, i64, vp, i64, vp]),
    "hoisdf_fold_weight_norm": (C.c_int, [vp, vp, i64, i64, vp, i64, vp, i64, vp]),
    "hoisdf_nchw_to_nhwc": (C.c_int, [vp, vp, i64, i64, i64, i64, vp]),
    "hoisdf_nchw_to_nhwc_split": (C.c_int, [vp, vp, vp, i64, i64, i64, i64, i64, vp]),
    "hoisdf_lattice_chunks": (C.c_int, [i32]),
    "hoisdf_lattice_count": (C.c_int, [vp, vp, vp, f32, i64, i32, vp, vp, vp]),
    "hoisdf_lattice_compact": (C.c_int, [vp, vp, vp, f32, i64, i32, vp, vp, vp, vp, vp]),
    "hoisdf_project_points": (C.c_int, [vp, vp, vp, f32, i64, i64, vp, vp, vp]),
    "hoisdf_gather_fwd": (C.c_int, [C.POINTER(Pyramid), vp, i64, vp, i64, i64, i32, vp, i32, vp, i64, vp]),
    "hoisdf_gather_split_fwd": (C.c_int, [C.POINTER(Pyramid), vp, i64, vp, i64, i64, i32, vp, i32, vp, vp, i64, vp]),
    "hoisdf_posenc_split_fwd": (C.c_int, [vp, vp, i64, i32, vp, vp, i64, vp]),
    "hoisdf_sdf_decoder_h3_fwd": (C.c_int, [C.POINTER(SdfWeightsH3), vp, vp, i64, i64, vp, vp, vp, vp, i64, vp, f32, vp]),
    "hoisdf_sdf_chain_fwd": (C.c_int, [C.POINTER(SdfChainArgs), vp]),
    "hoisdf_f32_to_f16": (C.c_int, [vp, vp, i64, vp]),
    "hoisdf_gather_sum_h16_fwd": (C.c_int, [C.POINTER(PyramidH), vp, i64, vp, i64, i64, vp, i32, vp, i64, vp]),
    "hoisdf_posenc_fwd": (C.c_int, [vp, vp, i64, i32, vp, i64, i64, vp]),
    "hoisdf_sdf_decoder_fwd": (C.c_int, [C.POINTER(SdfWeights), vp, i64, i64, vp, vp, vp, f32, vp]),
    "hoisdf_sdf_pad_input": (C.c_int, [vp, i64, vp, i64, vp]),
    "hoisdf_select_points": (C.c_int, [vp, vp, vp, i64, i64, i32, f32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "hoisdf_sdf_infer_keep": (i64, [i64, i64]),
    "hoisdf_sdf_infer_workspace_bytes": (i64, [i64, i64, i64, i64, i32]),
    "hoisdf_sdf_infer_plan": (C.c_int, [vp, vp, vp, f32, i64, i32, vp, vp, vp, vp]),
    "hoisdf_sdf_infer_fwd": (C.c_int, [C.POINTER(SdfInferArgs), vp]),
    "hoisdf_tokens_fwd": (C.c_int, [vp, vp, vp, i64, vp, vp, i64, i64, vp, i64, i64, vp]),
    "hoisdf_attention_workspace_bytes": (i64, [i64, i64, i64, i64]),
    "hoisdf_attention_fwd": (C.c_int, [vp, i64, vp, vp, i64, vp, i64, i64, i64, i64, i64, i64, vp, vp, i64, vp]),
    "hoisdf_attention_train_fwd": (C.c_int, [vp, i64, vp, vp, i64, vp, i64, vp, i64, i64, i64, i64, i64, f32, C.c_uint64, vp, i64, vp]),
    "hoisdf_attention_bwd_workspace_bytes": (i64, [i64, i64, i64, i64]),
    "hoisdf_attention_bwd": (C.c_int, [vp, i64, vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, f32,
                                       C.c_uint64, vp, i64, vp]),
    "hoisdf_attention_split_fwd": (C.c_int, [vp, i64, vp, vp, i64, vp, vp, i64, i64, i64, i64, i64, i64, vp, i64, vp]),
    "hoisdf_add_layernorm_fwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp]),
    "hoisdf_add_layernorm_split_fwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp, vp, i64, vp, vp, i64, vp]),
    "hoisdf_encoder_workspace_bytes": (i64, [i64, i64, i64, i32]),
    "hoisdf_encoder_fwd": (C.c_int, [C.POINTER(EncoderArgs), vp]),
    "hoisdf_decoder_workspace_bytes": (i64, [i64, i64, i64, i64, i32]),
    "hoisdf_decoder_fwd": (C.c_int, [C.POINTER(DecoderArgs), vp]),
    "hoisdf_vote_joints_fwd": (C.c_int, [vp, vp, vp, i64, i64, i64, vp, vp]),
    "hoisdf_mano_fwd": (C.c_int, [C.POINTER(ManoModel), vp, vp, i64, vp, vp, vp]),
    "hoisdf_mano_aa_fwd": (C.c_int, [C.POINTER(ManoModel), vp, vp, i64, vp, vp, vp]),
    "hoisdf_obj_metrics_workspace_bytes": (i64, [i64, i64]),
    "hoisdf_obj_metrics_fwd": (C.c_int, [vp, vp, i64, i64, vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp]),
    "hoisdf_mesh_metrics_fwd": (C.c_int, [vp, vp, i64, i64, vp, vp, vp, vp, i64, vp]),
    "hoisdf_hand_joint_metrics_fwd": (C.c_int, [vp, vp, i64, i64, vp, vp, vp, vp]),
    "hoisdf_gemm_f32": (C.c_int, [vp, i64, i32, vp, i64, i32, vp, i64, i64, i64, i64, i32, vp]),
    "hoisdf_gemm_f32_batched": (C.c_int, [vp, i64, i32, i64, i64, vp, i64, i32, i64, i64, vp, i64, i64, i64, i64, i64, i64,
                                          C.c_float, i32, i64, i64, vp]),
    "hoisdf_image_crop_fwd": (C.c_int, [vp, i64, i64, i64, i64, i64, i64, vp, vp, i64, C.c_float, vp, vp, vp, vp]),
    "hoisdf_gaussian_blur_params": (C.c_int, [C.c_float, i32, vp]),
    "hoisdf_gaussian_blur_u8": (C.c_int, [vp, vp, vp, i64, i64, i64, i64, vp, i32, vp]),
    "hoisdf_color_jitter_u8": (C.c_int, [vp, vp, i64, i64, i64, vp, vp, vp, vp]),
    "hoisdf_train_image_smem_bytes": (i64, [i64]),
    "hoisdf_train_image_fwd": (C.c_int, [vp, i64, i64, i64, i64, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp]),
    "hoisdf_mask_crop_fwd": (C.c_int, [vp, i64, i64, i64, i64, i64, vp, vp, i64, i64, vp, vp]),
    "hoisdf_sdf_rows_fwd": (C.c_int, [vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, vp, vp,
                                      vp, vp, vp, vp]),
    "hoisdf_linear_train_workspace_bytes": (i64, [i64, i64, i64]),
    "hoisdf_linear_train_fwd": (C.c_int, [vp, i64, vp, i64, vp, i64, i64, i64, i32, vp, i64, vp, i64, vp]),
    "hoisdf_linear_train_bwd": (C.c_int, [vp, i64, vp, i64, vp, i64, vp, i64, i64, i64, i64, i32, vp, i64, vp, i64, vp, vp,
                                          i64, vp]),
    "hoisdf_thin_linear_fwd": (C.c_int, [vp, i64, vp, i64, vp, i64, i64, i64, i32, vp, i64, vp]),
    "hoisdf_thin_linear_dw": (C.c_int, [vp, i64, vp, i64, i64, i64, i64, vp, i64, i32, vp]),
    "hoisdf_act_bias_bwd": (C.c_int, [vp, i64, vp, i64, i64, i64, i32, vp, i32, vp]),
    "hoisdf_weight_norm_bwd": (C.c_int, [vp, vp, vp, i64, i64, i64, vp, vp, i32, vp]),
    "hoisdf_gather_bwd": (C.c_int, [C.POINTER(Pyramid), vp, i64, vp, i64, i64, vp, i64, vp]),
    "hoisdf_sdf_loss_bwd": (C.c_int, [vp, vp, i64, f32, f32, vp, vp]),
    "hoisdf_layernorm_bwd": (C.c_int, [vp, vp, vp, i64, i64, vp, vp, vp, vp, i32, vp]),
    "hoisdf_softmax_rows_fwd": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, vp, i64, vp]),
    "hoisdf_softmax_rows_bwd": (C.c_int, [vp, i64, vp, i64, i64, i64, vp, i64, vp]),
    "hoisdf_absmax": (C.c_int, [vp, i64, i64, i64, vp, vp]),
    "hoisdf_linear_bwd_prep": (C.c_int, [vp, i64, vp, i64, i64, i64, i32, vp, vp, vp, i64, vp, vp, vp, i64, vp, vp, vp]),
    "hoisdf_split_rows_t": (C.c_int, [vp, i64, i64, i64, vp, vp, i64, vp]),
    "hoisdf_softmax_dropout_rows_fwd": (C.c_int, [vp, i64, i64, i64, i64, vp, i64, vp, i64, vp, i64, f32, C.c_uint64, vp]),
    "hoisdf_softmax_dropout_rows_bwd": (C.c_int, [vp, i64, vp, i64, i64, i64, vp, i64, f32, C.c_uint64, vp]),
    "hoisdf_tokens_bwd_workspace_bytes": (i64, [i64, i64]),
    "hoisdf_tokens_bwd": (C.c_int, [vp, i64, i64, vp, i64, vp, vp, i64, i64, vp, i64, vp, vp, i32, vp, i64, vp]),
    "hoisdf_vote_loss_bwd": (C.c_int, [vp, vp, vp, vp, i64, i64, i64, f32, f32, f32, f32, vp, vp, vp, vp]),
    "hoisdf_adamw_step": (C.c_int, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i64, vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "hoisdf_b200: %s is missing -- there is no CPU or PyTorch fallback; build the CUDA library with "
            "`python -m hoisdf_b200.csrc.build`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.hoisdf_abi_version() != ABI_VERSION:
        raise ImportError("hoisdf_b200: ABI mismatch (library %d, binding %d)" % (lib.hoisdf_abi_version(), ABI_VERSION))
    return lib


lib = _load()


E_UNSUPPORTED, E_WORKSPACE, E_TOO_FEW_POINTS = -4, -5, -6      # include/hoisdf_b200.h


class HoisdfError(RuntimeError):
    pass


def check(status: int, what: str):
    if status != 0:
        raise HoisdfError("%s failed: %s (status %d)" % (what, lib.hoisdf_status_string(status).decode(), status))
