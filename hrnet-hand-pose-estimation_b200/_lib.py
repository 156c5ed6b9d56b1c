"""ctypes binding of libhrnb.so (include/hrnb.h).

There is NO fallback: if the library cannot be loaded every product entry point raises.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("HRNB_LIB", "libhrnb.so"))

_lib = None
_lock = threading.Lock()

ABI_VERSION = 5            # include/hrnb.h HRNB_ABI_VERSION (tests/test_abi.py checks the header against this)
HRNB_CONV_RELU = 1
HRNB_CONV_OUT_NCHW = 2
HRNB_CONV_GATHER = 4
HRNB_CONV_IN_PHASES = 8
HRNB_CONV_OUT_PHASES = 16
HRNB_CONV_NO_PDL = 32
HRNB_CONV_FUSE_AFTER_RELU = 64


def guard_lead(Wp):
    return ((Wp + 2) + 7) // 8 * 8


def guard_tail(Wp):
    return ((Wp + 2) + 7) // 8 * 8 + 512


class ConvParams(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("in_ps", C.c_int64),
        ("wpk", C.c_void_p), ("bias", C.c_void_p),
        ("res", C.c_void_p), ("res_ps", C.c_int64),
        ("out", C.c_void_p), ("out_ps", C.c_int64),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("in_H", C.c_int32), ("in_W", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32), ("taps", C.c_int32), ("stride", C.c_int32),
        ("KC", C.c_int32), ("BN", C.c_int32), ("MB", C.c_int32), ("flags", C.c_int32),
        ("in_phase_stride", C.c_int64), ("out_phase_stride", C.c_int64),
        ("out2", C.c_void_p), ("out2_ps", C.c_int64), ("out2_phase_stride", C.c_int64),
        ("ntap_custom", C.c_int32), ("tap_src", C.c_int32 * 9), ("tap_dpos", C.c_int32 * 9),
        ("stats_sums", C.c_void_p), ("stats_ws", C.c_void_p),
        ("nfuse", C.c_int32), ("fuse_shift", C.c_int32 * 3), ("fuse_src", C.c_void_p * 3), ("fuse_ps", C.c_int64 * 3),
        ("in_up_shift", C.c_int32), ("ngroup", C.c_int32),
        ("grp_ntap", C.c_int32 * 4), ("grp_tap_dpos", (C.c_int32 * 4) * 4), ("grp_wpk", C.c_void_p * 4),
        ("grp_out_stride", C.c_int64), ("grp_res_stride", C.c_int64),
    ]


class FuseParams(C.Structure):
    _fields_ = [
        ("src", C.c_void_p * 4), ("src_ps", C.c_int64 * 4), ("shift", C.c_int32 * 4),
        ("nsrc", C.c_int32),
        ("out", C.c_void_p), ("out_ps", C.c_int64),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("relu", C.c_int32),
    ]


class WgradParams(C.Structure):
    _fields_ = [
        ("dy", C.c_void_p), ("dy_ps", C.c_int64), ("x", C.c_void_p), ("x_ps", C.c_int64), ("dw", C.c_void_p),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
        ("ntap", C.c_int32), ("tap_dpos", C.c_int32 * 9), ("tap_id", C.c_int32 * 9),
        ("NT", C.c_int32), ("TG", C.c_int32), ("KP", C.c_int32), ("ksplit", C.c_int32),
        ("tap_src", C.c_int32 * 9), ("pad_", C.c_int32), ("x_src_stride", C.c_int64),
    ]


class PackJob(C.Structure):
    _fields_ = [
        ("w", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("wpk_out", C.c_void_p), ("bias_out", C.c_void_p),
        ("cout", C.c_int32), ("cin", C.c_int32), ("taps_total", C.c_int32), ("transpose", C.c_int32),
        ("lcout", C.c_int32), ("lcin", C.c_int32), ("ntap", C.c_int32), ("tap_ids", C.c_int32 * 9),
        ("KC", C.c_int32), ("BN", C.c_int32), ("block0", C.c_int32), ("pad_", C.c_int32),
    ]


class BnParams(C.Structure):
    _fields_ = [
        ("c", C.c_void_p), ("c_ps", C.c_int64), ("sums", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("res", C.c_void_p), ("res_ps", C.c_int64), ("out", C.c_void_p), ("out_ps", C.c_int64),
        ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
        ("N", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("relu", C.c_int32),
        ("eps", C.c_float), ("momentum", C.c_float),
        ("out2", C.c_void_p), ("out2_ps", C.c_int64), ("out2_phase_stride", C.c_int64),
    ]


class BnBwdParams(C.Structure):
    _fields_ = [
        ("dy", C.c_void_p), ("dy_ps", C.c_int64), ("y", C.c_void_p), ("y_ps", C.c_int64),
        ("c", C.c_void_p), ("c_ps", C.c_int64), ("sums", C.c_void_p), ("gamma", C.c_void_p), ("dsums", C.c_void_p),
        ("ws", C.c_void_p), ("dc", C.c_void_p), ("dc_ps", C.c_int64), ("dres", C.c_void_p), ("dres_ps", C.c_int64), ("dres_mode", C.c_int32),
        ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
        ("N", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("relu", C.c_int32), ("eps", C.c_float),
    ]


class ParamSeg(C.Structure):
    _fields_ = [
        ("p_off", C.c_int64), ("g_off", C.c_int64), ("numel", C.c_int32),
        ("cout", C.c_int32), ("cin", C.c_int32), ("cin_g", C.c_int32), ("taps", C.c_int32),
        ("block0", C.c_int32), ("frozen", C.c_int32), ("pad_", C.c_int32),
    ]


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
_SIGS = {
    "hrnb_conv": (C.c_int, [C.POINTER(ConvParams), _vp]),
    "hrnb_conv_smem_bytes": (_i64, [C.POINTER(ConvParams)]),
    "hrnb_conv_stats_ws_floats": (_i64, []),
    "hrnb_pack_conv_weights": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hrnb_stem_conv1": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "hrnb_stem_im2col": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "hrnb_phase_split": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _i64, _vp]),
    "hrnb_fuse_sum": (C.c_int, [C.POINTER(FuseParams), _vp]),
    "hrnb_bilinear_up": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _i32, _vp]),
    "hrnb_pf8_to_nchw_f32": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "hrnb_nchw_f32_to_pf8": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "hrnb_decode_argmax": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "hrnb_softmax_softargmax": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hrnb_softargmax": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "hrnb_softmax_softargmax_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hrnb_final_preds": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
    "hrnb_loss_heatmap": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "hrnb_loss_pose2d": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "hrnb_wgrad": (C.c_int, [C.POINTER(WgradParams), _vp]),
    "hrnb_wgrad_smem_bytes": (_i64, [C.POINTER(WgradParams)]),
    "hrnb_pack_conv_weights_batch": (C.c_int, [_vp, _vp, _i32, _vp]),
    "hrnb_bn_stats": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hrnb_channel_sum": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hrnb_reduce_ws_floats": (_i64, []),
    "hrnb_bn_apply": (C.c_int, [C.POINTER(BnParams), _vp]),
    "hrnb_bn_forward_batch": (C.c_int, [C.POINTER(BnParams), _i32, _vp, _i32, _vp]),
    "hrnb_bn_backward_batch": (C.c_int, [C.POINTER(BnBwdParams), _i32, _vp]),
    "hrnb_bn_bwd_reduce": (C.c_int, [C.POINTER(BnBwdParams), _vp]),
    "hrnb_bn_bwd_apply": (C.c_int, [C.POINTER(BnBwdParams), _vp]),
    "hrnb_fuse_sum_bwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "hrnb_fuse_sum_bwd_batch": (C.c_int, [_vp, _i64, _vp, _i64, _i32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i32, _i32, _i32, _i32, _i32, _vp]),
    "hrnb_bilinear_up_bwd": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "hrnb_phase_merge": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "hrnb_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "hrnb_adam_tick": (C.c_int, [_vp, _vp, _vp]),
    "hrnb_grad_to_natural": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "hrnb_triangulate_dlt": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "hrnb_triangulate_dlt_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "hrnb_gen_heatmaps": (C.c_int, [_vp, _i32, _i32, _i32, _i32, C.c_float, _vp, _vp]),
    "hrnb_stem_im2col_u8": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), _vp, _i64, _i32, _i32, _i32, _vp]),
    "hrnb_flip_merge": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "hrnb_maxpool2_relu": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "hrnb_gap_mlp": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp]),
    "hrnb_axpby": (C.c_int, [C.c_float, _vp, C.c_float, _vp, _vp, _i64, _vp]),
    "hrnb_last_error": (C.c_char_p, []),
    "hrnb_abi_version": (C.c_int, []),
    "hrnb_launch_count": (_i64, []),
    "hrnb_hang_init": (C.c_int, []),
    "hrnb_hang_report": (C.c_int, [_vp, C.c_int]),
    "hrnb_debug_set": (C.c_int, [C.c_int, C.c_int]),
    "hrnb_debug_trace": (C.c_int, [_vp]),
}
EXPORTS = tuple(_SIGS)


class HrnbError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if libhrnb.so is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise HrnbError(
                "libhrnb.so not found at %s - run `python hrnet-hand-pose-estimation_b200/build.py` "
                "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            if os.environ.get("HRNB_LIB_LAX", "0") == "1" and not hasattr(h, name):
                continue           # A/B runs against libraries built from earlier commits (tools/gpu_trip21.sh)
            fn = getattr(h, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if h.hrnb_abi_version() != ABI_VERSION:
            raise HrnbError("libhrnb.so ABI version mismatch")
        if os.environ.get("HRNB_NO_PDL", "0") == "1":      # debug: launch the conv kernels without programmatic dependent launch
            h.hrnb_debug_set(2, 1)
        if os.environ.get("HRNB_TWIN_MIN"):                 # two-CTAs-per-SM variant: tiles per CTA slot it needs (huge = off)
            h.hrnb_debug_set(9, int(os.environ["HRNB_TWIN_MIN"]))
        if os.environ.get("HRNB_SLAB", "0") == "1":        # opt-in: resident multi-chunk / multi-N-tile weight slabs (ops.py too)
            h.hrnb_debug_set(10, 1)
        if os.environ.get("HRNB_NO_DUAL", "0") == "1":     # A/B: one MMA-issuing warp per CTA everywhere
            h.hrnb_debug_set(8, 1)
        if os.environ.get("HRNB_TMEM_SHARE", "0") == "1":  # debug: let TMEM-holding CTAs of different kernels share an SM (can deadlock)
            h.hrnb_debug_set(6, 1)
        _lib = h
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().hrnb_last_error()
        raise HrnbError("libhrnb call failed (%d): %s" % (rc, msg.decode() if msg else "?"))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def hang_init():
    """arm the mbarrier time-out records on the current device (engines call this once; not under stream capture)"""
    check(lib().hrnb_hang_init())


def hang_report():
    """-> [] or the records the stuck warps left before the trap: dicts(kernel, grid, cta, warp, barrier, parity)"""
    buf = (C.c_uint64 * 512)()
    n = lib().hrnb_hang_report(buf, 512)
    if n < 2 or buf[0] == 0:
        return []
    out = []
    for i in range(2, n - 1, 2):
        w0, w1 = buf[i], buf[i + 1]
        if w0 == 0 and w1 == 0:
            continue
        threads = w0 >> 48
        out.append({"kernel": {608: "conv_tc", 736: "conv_tc<gather>", 192: "wgrad_tc"}.get(threads, "threads=%d" % threads),
                    "grid": (w0 >> 32) & 0xffff, "cta": (w0 >> 8) & 0xffffff, "warp": w0 & 0xff,
                    "barrier_smem": w1 >> 8, "parity": w1 & 0xff})
    return out


def launch_count():
    return int(lib().hrnb_launch_count())
