"""ctypes binding of the C ABI declared in include/mde_b200.h (libmde_b200.so, built by __graft_entry__.build()).

The library is loaded lazily on first use.  There is NO fallback: if the shared object is missing, or the
process has no sm_100 GPU, the first operator call raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmde_b200.so")

_i32, _i64, _f32 = ctypes.c_int, ctypes.c_int64, ctypes.c_float
_p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/mde_b200.h one to one
SIGNATURES = {
    "mde_version": (_i32, []),
    "mde_error_string": (ctypes.c_char_p, [_i32]),
    "mde_check_device": (_i32, []),
    "mde_launch_count": (_i64, []),
    "mde_set_pdl": (_i32, [_i32]),
    "mde_gather_embed": (_i32, [_p, _p, _p, _p, _i32, _i64, _i32, _i32, _i32, _i32, _i64, _p, _p]),
    "mde_gather_embed_labels": (_i32, [_p, _i32, _p, _p, _p, _i32, _i64, _i32, _i32, _i32, _i32, _i64, _p, _p]),
    "mde_gather_embed_nhwc": (_i32, [_p, _i32, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_class_area_table": (_i32, [_p, _i32, _i64, _i32, _p, _p, _p]),
    "mde_cast_i64_f32": (_i32, [_p, _p, _i64, _p]),
    "mde_aux_mlp_fwd": (_i32, [_p, _i64, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i64, _f32, _p]),
    "mde_aux_mlp_bwd": (_i32, [_p, _i64, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i64,
                               _f32, _p]),
    "mde_bias_act_nhwc": (_i32, [_p, _p, _p, _p, _i64, _i32, _i32, _p]),
    "mde_bias_act_pad_nhwc": (_i32, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_regressor_bins_fwd": (_i32, [_p, _i64, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _f32, _f32,
                                      _p, _p, _p, _p, _p]),
    "mde_patch_embed_ws_floats": (_i64, [_i32, _i32, _i32, _i32, _i32]),
    "mde_patch_embed_fwd": (_i32, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_conv3x3_prep_weight": (_i32, [_p, _p, _i32, _i32, _f32, _p]),
    "mde_conv3x3_nhwc_fwd": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _p]),
    "mde_conv3x3_prep_weight_x3": (_i32, [_p, _p, _i32, _i32, _p]),
    "mde_conv3x3_nhwc_x3_fwd": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _p]),
    "mde_conv3x3_small_nhwc_fwd": (_i32, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_split_bf16": (_i32, [_p, _p, _i64, _p]),
    "mde_merge_bf16": (_i32, [_p, _p, _i64, _p]),
    "mde_split_bf16_nchw": (_i32, [_p, _p, _i32, _i32, _i64, _p]),
    "mde_range_attention_tc": (_i32, [_p, _p, _p, _i32, _i32, _i32, _i64, _p]),
    "mde_upsample_concat_nhwc_pair_fwd": (_i32, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_pointwise_x3_fwd": (_i32, [_p, _p, _i64, _p, _p, _i32, _p, _p, _i64, _i32, _i32, _i64, _i64, _p, _p]),
    "mde_pool_slabs": (_i32, [_i32, _i64]),
    "mde_stem_conv3x3s2_nhwc": (_i32, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_depthwise_bias_act_pool_nhwc": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                                _i32, _p]),
    "mde_bias_act_pool_nhwc": (_i32, [_p, _p, _p, _p, _i32, _i64, _i32, _i32, _p]),
    "mde_se_gate": (_i32, [_p, _i32, _f32, _p, _p, _p, _p, _p, _i32, _i32, _i32, _p]),
    "mde_gemm_nt_tf32": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _p]),
    "mde_gemm_nt_tf32_ex": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _p, _i32, _i32, _p]),
    "mde_gemm_nt_tf32_planes": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _p, _i32, _i32, _i64, _p]),
    "mde_nhwc_to_cpad_tf32": (_i32, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_conv3x3_wgrad_tf32": (_i32, [_p, _p, _p, _i32, _i32, _i64, _i64, _i32, _i32, _p]),
    "mde_linear_fwd": (_i32, [_p, _i32, _p, _i32, _p, _p, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_bins_finalize_fwd": (_i32, [_p, _i32, _i32, _i32, _f32, _f32, _p, _p, _p, _p]),
    "mde_encoder_layer_ws_floats": (_i64, [_i32, _i32, _i32, _i32]),
    "mde_encoder_layer_fwd": (_i32, [_p] * 15 + [_i32, _i32, _i32, _i32, _i32, _f32, _p]),
    "mde_encoder_layer_tc_ws_floats": (_i64, [_i32, _i32, _i32, _i32]),
    "mde_split3_tf32": (_i32, [_p, _p, _i64, _i32, _p]),
    "mde_encoder_layer_tc_fwd": (_i32, [_p, _p, _i32] + [_p] * 13 + [_i32, _i32, _i32, _i32, _i32, _f32, _p]),
    "mde_range_attention": (_i32, [_p, _p, _p, _i32, _i32, _i32, _i64, _i32, _p]),
    "mde_bins_pred_fwd": (_i32, [_p, _p, _p, _i32, _i32, _i64, _p]),
    "mde_conv1x1_fwd": (_i32, [_p, _p, _p, _p, _i32, _i32, _i32, _i64, _p]),
    "mde_head_chain_fwd": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i64, _p]),
    "mde_head_chain_bf16_fwd": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i64, _p]),
    "mde_head_chain_fwd_train": (_i32, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i64, _p]),
    "mde_head_chain_bwd_logits": (_i32, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i64, _p]),
    "mde_fold_queries": (_i32, [_p, _p, _p, _i64, _p, _p, _p, _i32, _i32, _i32, _i32, _f32, _i32, _p]),
    "mde_round_tf32": (_i32, [_p, _p, _i64, _f32, _p]),
    "mde_tc_last_error": (_i32, []),
    "mde_tc_debug_profile": (_i32, [_p]),
    "mde_upsample_concat_fwd": (_i32, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_upsample_concat_nhwc_fwd": (_i32, [_p, _p, _i32, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_upsample_bwd_ws_bytes": (_i64, [_i32, _i32]),
    "mde_upsample_bwd": (_i32, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p]),
    "mde_nchw_to_nhwc_slice": (_i32, [_p, _p, _i32, _i32, _i64, _i32, _p]),
    "mde_nchw_to_nhwc_slice_padded": (_i32, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mde_nchw_to_nhwc": (_i32, [_p, _p, _i32, _i32, _i64, _p]),
    "mde_relu_eps_fwd": (_i32, [_p, _p, _i64, _f32, _p]),
    "mde_eval_metrics_ws_bytes": (_i64, [_i32]),
    "mde_eval_metrics_fwd": (_i32, [_p, _p, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "mde_flip_average": (_i32, [_p, _p, _p, _i64, _i32, _f32, _f32, _p]),
    "mde_bn_stats_nhwc": (_i32, [_p, _i64, _i32, _p, _p, _p]),
    "mde_bn_apply_nhwc": (_i32, [_p, _p, _i64, _i32, _p, ctypes.c_double, _p, _p, _f32, _p, _p, _p, _p, _f32, _p]),
    "mde_bn_bwd_reduce_nhwc": (_i32, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p]),
    "mde_bn_bwd_apply_nhwc": (_i32, [_p, _p, _p, _i64, _i32, _p, _p, _p, _p, ctypes.c_double, _p]),
    "mde_bn_stats_p2p_nhwc": (_i32, [_p, _i64, _i32, _p, _p, _p, _i32, _i32, _i64, _i64, ctypes.c_uint64, _p]),
    "mde_bn_apply_p2p_nhwc": (_i32, [_p, _p, _i64, _i32, ctypes.c_uint64, _i64, _i64, _i32, ctypes.c_uint64, ctypes.c_double,
                                     _p, _p, _f32, _p, _p, _p, _p, _f32, _p]),
    "mde_bn_bwd_reduce_p2p_nhwc": (_i32, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p, _i32, _i32, _i64, _i64, ctypes.c_uint64, _p]),
    "mde_bn_bwd_apply_p2p_nhwc": (_i32, [_p, _p, _p, _i64, _i32, _p, _p, _p, ctypes.c_uint64, _i64, _i64, _i32,
                                         ctypes.c_uint64, ctypes.c_double, _p]),
    "mde_bn_p2p_set_epoch_counter": (_i32, [_p]),
    "mde_bn_set_peer_timeout_seconds": (_i32, [ctypes.c_double]),
    "mde_bn_peer_timeouts": (_i32, []),
    "mde_bn_wait_stats": (_i32, [_p, _p, _i32]),
    "mde_silog_ws_bytes": (_i64, []),
    "mde_silog_fwd": (_i32, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "mde_silog_bwd": (_i32, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p]),
    "mde_chamfer_ws_bytes": (_i64, [_i32, _i32]),
    "mde_chamfer_fwd": (_i32, [_p, _p, _i32, _i32, _i64, _f32, _i32, _p, _p, _p]),
    "mde_depth_losses_fwd": (_i32, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _i32, _p, _p, _p, _p, _p]),
    "mde_silog_bwd_thr": (_i32, [_p, _p, _f32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p]),
    "mde_chamfer_bwd": (_i32, [_p, _i32, _i32, _p, _p, _p, _p]),
}

_lib = None
_lock = threading.Lock()


class MdeError(RuntimeError):
    pass


def load(check_device=True):
    """Load libmde_b200.so (once).  Raises MdeError if it is missing or, when check_device, if there is no B200."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise MdeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                    "There is no CPU or PyTorch fallback for the hot path.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError here == header/library mismatch
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    if check_device:
        rc = _lib.mde_check_device()
        if rc != 0:
            raise MdeError("mde_b200 kernels need an sm_100 (B200) device: " + _lib.mde_error_string(rc).decode())
    return _lib


def check(rc, what):
    if rc != 0:
        raise MdeError(f"{what} failed: {load(False).mde_error_string(rc).decode()} ({rc})")
