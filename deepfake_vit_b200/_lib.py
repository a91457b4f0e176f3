"""ctypes binding of libdfvit.so (C ABI declared in include/dfvit.h).

There is no fallback: if the shared library is missing, importing this module raises, and
every launch entry point refuses non-sm_100 devices (DFV_ERR_DEVICE) -- the product has no
CPU or stock-PyTorch code path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdfvit.so")

DFV_F32, DFV_BF16 = 0, 1
DFV_ACT_NONE, DFV_ACT_SILU, DFV_ACT_RELU = 0, 1, 2
(W_STEM, W_STEM_BIAS, W_EXPAND, W_EXPAND_BIAS, W_DW, W_DW_BIAS, W_SE_REDUCE, W_SE_REDUCE_BIAS,
 W_SE_EXPAND, W_SE_EXPAND_BIAS, W_PROJECT, W_PROJECT_BIAS, W_HEAD, W_HEAD_BIAS) = range(14)


class DfvError(RuntimeError):
    pass


class BlockInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "c_in", "c_mid", "c_out", "kernel", "stride", "pad_lo", "pad_hi", "se_squeeze", "has_expand", "has_skip")]


class InferArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("use_attention", C.c_int32), ("use_landmark", C.c_int32), ("use_channel", C.c_int32),
        ("use_spatial", C.c_int32), ("heat_group", C.c_int32), ("landmark_ref_size", C.c_float),
        ("blob", C.c_void_p), ("images_nchw", C.c_void_p), ("landmarks", C.c_void_p),
        ("lm_weights", C.c_void_p), ("ca_w1", C.c_void_p), ("ca_w2_t", C.c_void_p),
        ("ca_hidden", C.c_int32), ("sa_w", C.c_void_p),
        ("head_w_t", C.POINTER(C.c_void_p)), ("head_b", C.POINTER(C.c_void_p)),
        ("head_dims", C.POINTER(C.c_int32)), ("head_layers", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("logits", C.c_void_p), ("features", C.c_void_p), ("heat", C.c_void_p),
        ("taps", C.POINTER(C.c_void_p)),
        ("images_u8", C.c_void_p), ("u8_norm", C.c_float * 6),
        ("heat_max_floor", C.c_void_p),
    ]


class GemmTuning(C.Structure):
    _fields_ = [("weight_stationary", C.c_int32), ("bn", C.c_int32), ("cluster", C.c_int32), ("share_a", C.c_int32), ("rotate", C.c_int32)]


class DwconvTuning(C.Structure):
    _fields_ = [("L", C.c_int32), ("TW", C.c_int32), ("TH", C.c_int32), ("CB", C.c_int32)]


class PackArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("bn_eps", C.c_float), ("cls_bn_eps", C.c_float),
        ("params", C.POINTER(C.c_void_p)), ("blob", C.c_void_p),
        ("head_layers", C.c_int32), ("head_dims", C.POINTER(C.c_int32)),
        ("head_w_t", C.POINTER(C.c_void_p)), ("head_b", C.POINTER(C.c_void_p)),
        ("ca_hidden", C.c_int32), ("ca_w2_t", C.c_void_p),
    ]


GRAD_UNITS = 34     # include/dfvit.h DFV_GRAD_UNITS: classifier+attention+head conv, block 31 .. block 0, stem


class TrainArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("use_attention", C.c_int32), ("use_landmark", C.c_int32), ("use_channel", C.c_int32),
        ("use_spatial", C.c_int32), ("heat_group", C.c_int32), ("landmark_ref_size", C.c_float),
        ("bn_eps", C.c_float), ("bn_momentum", C.c_float), ("cls_bn_eps", C.c_float), ("cls_bn_momentum", C.c_float),
        ("drop_connect_rate", C.c_float), ("feat_dropout", C.c_float), ("cls_dropout", C.c_float),
        ("seed", C.c_uint64),
        ("params", C.POINTER(C.c_void_p)), ("grads", C.POINTER(C.c_void_p)),
        ("images_nchw", C.c_void_p), ("landmarks", C.c_void_p),
        ("ca_hidden", C.c_int32), ("head_dims", C.POINTER(C.c_int32)), ("head_layers", C.c_int32),
        ("arena", C.c_void_p), ("arena_bytes", C.c_size_t), ("scratch", C.c_void_p), ("scratch_bytes", C.c_size_t),
        ("logits", C.c_void_p), ("features", C.c_void_p), ("dlogits", C.c_void_p), ("dfeatures", C.c_void_p),
        ("taps", C.POINTER(C.c_void_p)),
        ("freeze_bn", C.c_int32), ("grad_events", C.POINTER(C.c_void_p)), ("seed_dev", C.c_void_p),
    ]


# per-block / global tensor kinds of the training parameter table (include/dfvit.h DFV_T_*, DFV_TG_*)
(T_EXPAND_W, T_BN0_G, T_BN0_B, T_BN0_RM, T_BN0_RV, T_DW_W, T_BN1_G, T_BN1_B, T_BN1_RM, T_BN1_RV,
 T_SE_R_W, T_SE_R_B, T_SE_E_W, T_SE_E_B, T_PROJ_W, T_BN2_G, T_BN2_B, T_BN2_RM, T_BN2_RV) = range(19)
(TG_STEM_W, TG_STEM_G, TG_STEM_B, TG_STEM_RM, TG_STEM_RV, TG_HEAD_W, TG_HEAD_G, TG_HEAD_B, TG_HEAD_RM, TG_HEAD_RV,
 TG_LM_W, TG_SA_W, TG_CA_W1, TG_CA_W2) = range(14)


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C deepfake_vit_b200/csrc`). deepfake_vit_b200 has no fallback path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
    sigs = {
        "dfv_version": (C.c_int, []),
        "dfv_last_error": (C.c_char_p, []),
        "dfv_device_check": (C.c_int, []),
        "dfv_launch_count": (i64, [i32]),
        "dfv_last_timeout_word": (C.c_uint, []),
        "dfv_profile_enable": (C.c_int, [i32]),
        "dfv_profile_count": (C.c_int, []),
        "dfv_profile_get": (C.c_int, [i32, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_float)]),
        "dfv_b4_num_blocks": (C.c_int, []),
        "dfv_b4_block": (C.c_int, [i32, C.POINTER(BlockInfo)]),
        "dfv_b4_stem_channels": (C.c_int, []),
        "dfv_b4_head_channels": (C.c_int, []),
        "dfv_b4_output_hw": (C.c_int, [i32, i32, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "dfv_blob_bytes": (sz, [i32]),
        "dfv_blob_slot": (C.c_int, [i32, i32, i32, C.POINTER(sz), C.POINTER(sz)]),
        "dfv_stem_conv_fwd": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
        "dfv_stem_conv_u8_fwd": (C.c_int, [vp, C.POINTER(C.c_float), vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
        "dfv_u8_to_nchw_f32": (C.c_int, [vp, C.POINTER(C.c_float), vp, i32, i32, i32, vp]),
        "dfv_pack_weights": (C.c_int, [C.POINTER(PackArgs), vp]),
        "dfv_pw_gemm_fwd_tuned": (C.c_int, [vp, vp, vp, vp, i32, vp, vp, i32, i64, i32, i32, i32, C.POINTER(GemmTuning), vp]),
        "dfv_dwconv_fwd_tuned": (C.c_int, [vp, vp, vp, vp, vp] + [i32] * 10 + [C.POINTER(DwconvTuning), vp]),
        "dfv_dwconv_pool_parts_tuned": (C.c_int, [i32] * 9 + [C.POINTER(DwconvTuning)]),
        "dfv_dwconv_pool_parts": (C.c_int, [i32] * 9),
        "dfv_dwconv_fwd": (C.c_int, [vp, vp, vp, vp, vp] + [i32] * 10 + [vp]),
        "dfv_dwconv_se_supported": (C.c_int, [i32] * 10),
        "dfv_dwconv_se_profitable": (C.c_int, [i32] * 10),
        "dfv_dwconv_se_fwd": (C.c_int, [vp] * 8 + [i64] + [i32] * 10 + [vp]),
        "dfv_se_excite_fwd": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
        "dfv_se_scratch_floats": (sz, [i32, i32, i32]),
        "dfv_se_gate_fwd": (C.c_int, [vp, i32, f32, vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, vp]),
        "dfv_pw_gemm_fwd": (C.c_int, [vp, vp, vp, vp, i32, vp, vp, i32, i64, i32, i32, i32, vp]),
        "dfv_pw_conv_fwd": (C.c_int, [vp, vp, vp, vp, i32, vp, vp, i32, i32, i64, i32, i32, i32, vp, vp]),
        "dfv_pw_fold_ws_bytes": (sz, [i32]),
        "dfv_dwconv_stats_fwd": (C.c_int, [vp, vp, vp, vp, vp] + [i32] * 9 + [vp]),
        "dfv_bn_stats_from_sums": (C.c_int, [vp, i32, C.c_double, f32, f32, vp, vp, vp, vp, vp]),
        "dfv_dwconv_plan_info": (C.c_int, [i32] * 9 + [C.POINTER(C.c_int)]),
        "dfv_gemm_plan_info": (C.c_int, [C.c_longlong, i32, i32, i32, C.POINTER(C.c_int)]),
        "dfv_clip_aggregate": (C.c_int, [vp, i32, i32, i32, vp, vp, vp, f32, vp]),
        "dfv_global_avg_pool": (C.c_int, [vp, i32, vp, i32, i64, i32, vp]),
        "dfv_l2_normalize": (C.c_int, [vp, vp, i32, i32, f32, vp]),
        "dfv_clip_adamw_step": (C.c_int, [vp, vp, vp, vp, i64, vp] + [C.c_double] * 7 + [i64, vp, vp, vp]),
        "dfv_landmark_heatmap_fwd": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, i32, vp]),
        "dfv_landmark_heatmap_fwd_ex": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, i32, vp, vp]),
        "dfv_class_weight_sum": (C.c_int, [vp, vp, vp, i32, i32, vp]),
        "dfv_attention_scratch_floats": (sz, [i32] * 5),
        "dfv_hybrid_attention_fwd": (C.c_int, [vp] * 9 + [i32] * 8 + [vp]),
        "dfv_mlp_head_fwd": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int), i32, vp, vp, i32, vp]),
        "dfv_mlp_head_scratch_floats": (sz, [C.POINTER(C.c_int), i32, i32]),
        "dfv_combined_loss_fwd_bwd": (C.c_int, [vp, vp, vp, vp, f32, f32, f32, vp, vp, vp, i32, i32, i32,
                                                C.POINTER(C.c_int), vp]),
        "dfv_combined_loss_fwd_bwd_ex": (C.c_int, [vp, vp, vp, vp, f32, f32, f32, vp, vp, vp, i32, i32, i32,
                                                   C.POINTER(C.c_int), vp, vp]),
        "dfv_rows_chunks": (C.c_int, [i32, i64]),
        "dfv_bn_ws_floats": (sz, [i32, i64, i32]),
        "dfv_bn_stats_fwd": (C.c_int, [vp, i32, i32, i64, i32, f32, f32, vp, vp, vp, vp, vp, vp]),
        "dfv_bn_act_fwd": (C.c_int, [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, i64, i32, vp]),
        "dfv_act_bn_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i64,
                                     i32, vp]),
        "dfv_bn_bwd_apply": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, i32, i64, i32, vp]),
        "dfv_act_bn_bwd_apply": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, vp, vp, f32, vp, vp, vp, vp, i32, i32, i64, i32, vp]),
        "dfv_se_train_fwd": (C.c_int, [vp, i32, f32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, vp]),
        "dfv_se_bwd_ws_floats": (sz, [i32, i64, i32, i32]),
        "dfv_se_bwd": (C.c_int, [vp, vp, i32] + [vp] * 11 + [i32, i64, i32, i32, vp]),
        "dfv_linear_f32_scratch_floats": (sz, [i32, i32, i32]),
        "dfv_linear_f32_fwd": (C.c_int, [vp, vp, vp, vp, vp, sz, i32, i32, i32, i32, vp]),
        "dfv_se_bwd_from_partials": (C.c_int, [vp] * 11 + [i32, i64, i32, i32, vp]),
        "dfv_act_bn_bwd_gated_reduce": (C.c_int, [vp] * 8 + [i32, i32, i64, i32, vp]),
        "dfv_bn_bwd_gated_finalize": (C.c_int, [vp, vp, vp, f32, vp, vp, vp, vp, vp, i32, i32, i64, i32, vp]),
        "dfv_pw_wgrad": (C.c_int, [vp, vp, vp, i32, vp, i32, i64, i32, i32, vp]),
        "dfv_dwconv_dgrad": (C.c_int, [vp, vp, vp] + [i32] * 9 + [vp]),
        "dfv_dwconv_wgrad": (C.c_int, [vp, vp, vp] + [i32] * 9 + [vp]),
        "dfv_stem_wgrad": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, vp]),
        "dfv_attention_saved_floats": (sz, [i32] * 5),
        "dfv_hybrid_attention_train_fwd": (C.c_int, [vp] * 7 + [i32] * 8 + [vp]),
        "dfv_hybrid_attention_bwd": (C.c_int, [vp] * 13 + [i32] * 8 + [vp]),
        "dfv_landmark_heatmap_bwd": (C.c_int, [vp] * 6 + [i32, i32, i32, f32, f32, i32, vp]),
        "dfv_cast_weight": (C.c_int, [vp, vp, i32, i32, i32, i32, vp]),
        "dfv_dw_weight_pack": (C.c_int, [vp, vp, i32, i32, i32, vp]),
        "dfv_dw_weight_unpack": (C.c_int, [vp, vp, i32, i32, vp]),
        "dfv_dropout_mask": (C.c_int, [vp, i64, f32, C.c_uint64, vp]),
        "dfv_dropout_mask_dev": (C.c_int, [vp, i64, f32, C.c_uint64, vp, vp]),
        "dfv_colsum": (C.c_int, [vp, i32, i32, vp, vp]),
        "dfv_add_mul": (C.c_int, [vp, vp, vp, vp, i64, vp]),
        "dfv_convert": (C.c_int, [vp, i32, vp, i32, i64, vp]),
        "dfv_train_table_size": (C.c_int, []),
        "dfv_train_index": (C.c_int, [i32, i32]),
        "dfv_train_cls_index": (C.c_int, [i32, i32]),
        "dfv_train_arena_bytes": (sz, [i32, i32, i32, i32, C.POINTER(C.c_int32), i32, i32]),
        "dfv_train_scratch_bytes": (sz, [i32, i32, i32, i32, C.POINTER(C.c_int32), i32, i32]),
        "dfv_train_fwd": (C.c_int, [C.POINTER(TrainArgs), vp]),
        "dfv_train_bwd": (C.c_int, [C.POINTER(TrainArgs), vp]),
        "dfv_infer_workspace_bytes": (sz, [i32, i32, i32, i32]),
        "dfv_infer_fwd": (C.c_int, [C.POINTER(InferArgs), vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)          # AttributeError if the header and library disagree
        fn.restype, fn.argtypes = res, args
    return lib, tuple(sigs)


lib, EXPORTS = _load()


def check(rc: int):
    if rc != 0:
        raise DfvError(f"libdfvit error {rc}: {lib.dfv_last_error().decode(errors='replace')} "
                       f"[timeout word 0x{lib.dfv_last_timeout_word():08x}]")


PROFILE_KINDS = ("stem", "expand_gemm", "dwconv", "se_gate", "project_gemm", "heatmap", "attention", "mlp_head",
                 "loss", "gemm_simt", "bn_act", "pw_wgrad", "dwconv_bwd")


def profile_records():
    """[(kind_name, algorithmic_bytes, flops, ms)] for every launch recorded since dfv_profile_enable(1)."""
    out = []
    k, b, f, ms = C.c_int(), C.c_double(), C.c_double(), C.c_float()
    for i in range(lib.dfv_profile_count()):
        check(lib.dfv_profile_get(i, C.byref(k), C.byref(b), C.byref(f), C.byref(ms)))
        out.append((PROFILE_KINDS[k.value], b.value, f.value, ms.value))
    return out


def b4_blocks():
    out = []
    for i in range(lib.dfv_b4_num_blocks()):
        bi = BlockInfo()
        check(lib.dfv_b4_block(i, C.byref(bi)))
        out.append(bi)
    return out


def blob_slot(dtype: int, block: int, kind: int):
    off, n = C.c_size_t(), C.c_size_t()
    check(lib.dfv_blob_slot(dtype, block, kind, C.byref(off), C.byref(n)))
    return off.value, n.value
