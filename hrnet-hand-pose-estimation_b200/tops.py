"""Python wrappers over the TRAINING entry points of the C ABI (include/hrnb.h, "training path"): one call = one
kernel launch on torch's current stream.  Also the tap tables that express the data- and weight-gradients of the
3x3 / stride-2 convolutions as flat-shift GEMMs on PF8 tensors."""
import ctypes as C

import torch

from . import _lib
from ._lib import BnBwdParams, BnParams, WgradParams
from .pf8 import PF8, PhasePF8

EPS = 1e-5
MOMENTUM = 0.1


# ---- tap tables -------------------------------------------------------------------------------------------
def fwd_taps_s1(k, Wp):
    """[(dpos, tap_id)] of a stride-1 conv on a grid with padded width Wp: tap (r,s) reads p + (r-1)*Wp + (s-1)."""
    if k == 1:
        return [(0, 0)]
    return [((r - 1) * Wp + (s - 1), r * 3 + s) for r in range(3) for s in range(3)]


def fwd_taps_s2(Wp_half):
    """3x3 stride-2 pad-1 conv over a phase-split input: {phase: [(dpos, tap_id)]} on the half-resolution grid.
    Tap (r,s) reads phase (r != 1, s != 1) at (dy,dx) = (-(r == 0), -(s == 0))  (include/hrnb.h HRNB_CONV_IN_PHASES)."""
    out = {}
    for r in range(3):
        for s in range(3):
            ph = (r != 1) * 2 + (s != 1)
            out.setdefault(ph, []).append((-(r == 0) * Wp_half - (s == 0), r * 3 + s))
    return out


def dgrad_taps_s1(k, Wp):
    """data-gradient of a stride-1 conv = conv with transposed weights and mirrored taps:
    -> (tap_ids for packing, None = standard geometry)"""
    return ([0], None) if k == 1 else ([8 - t for t in range(9)], None)


def dgrad_taps_s2(Wp_half):
    """data-gradient of the 3x3 stride-2 conv, one flat-shift conv per INPUT phase on the half-resolution grid:
    {phase: (tap_ids, [(src, dpos)])}: d_phase[q] = sum_taps W_tap^T dc[q - dpos_fwd(tap)]."""
    out = {}
    for ph, taps in fwd_taps_s2(Wp_half).items():
        out[ph] = ([tid for _, tid in taps], [(0, -dpos) for dpos, _ in taps])
    return out


# ---- wgrad ------------------------------------------------------------------------------------------------
def wgrad_params(dy, x_ptr, x_ps, dw, cin, cout, taps, NT=0, TG=0, KP=0, ksplit=0, src_stride=0):
    """dy: PF8 gradient on the conv's output grid; x_ptr/x_ps: PF8 input tensor (or one phase of it) on the same grid;
    dw: fp32 [taps_total, cin, cout] (accumulated); taps: [(dpos, tap_id)] or, with src_stride (elements between the input
    tensors, e.g. the phases of a PhasePF8), [(dpos, tap_id, source)] grouped by source."""
    p = WgradParams()
    p.dy, p.dy_ps, p.x, p.x_ps, p.dw = dy.ptr, dy.ps, x_ptr, x_ps, dw.data_ptr()
    p.N, p.H, p.W, p.cin, p.cout, p.ntap = dy.N, dy.H, dy.W, cin, cout, len(taps)
    for t, tap in enumerate(taps):
        p.tap_dpos[t], p.tap_id[t] = tap[0], tap[1]
        p.tap_src[t] = tap[2] if len(tap) > 2 else 0
    p.NT, p.TG, p.KP, p.ksplit, p.x_src_stride = NT, TG, KP, ksplit, src_stride
    return p


def fwd_taps_s2_merged(Wp_half):
    """all nine taps of the 3x3 stride-2 conv over a phase-split input for ONE weight-gradient launch:
    [(dpos, tap_id, phase)] grouped by phase"""
    return [(dpos, tid, ph) for ph, taps in sorted(fwd_taps_s2(Wp_half).items()) for dpos, tid in taps]


def wgrad(dy, x, dw, cin, cout, taps, **kw):
    p = wgrad_params(dy, x.ptr, x.ps, dw, cin, cout, taps, **kw)
    _lib.check(_lib.lib().hrnb_wgrad(C.byref(p), _lib.stream_ptr()))
    return dw


def wgrad_conv(dy, x, dw, k, stride):
    """Weight gradient of a whole conv (1x1 / 3x3, stride 1 or 2 over a PhasePF8 input) into dw [k*k, cin, cout]."""
    cin, cout = dw.shape[1], dw.shape[2]
    if stride == 1:
        wgrad(dy, x, dw, cin, cout, fwd_taps_s1(k, dy.Wp))
    else:
        assert isinstance(x, PhasePF8) and k == 3
        p = wgrad_params(dy, x.ptr, x.ps, dw, cin, cout, fwd_taps_s2_merged(dy.Wp), src_stride=x.phase_stride)
        _lib.check(_lib.lib().hrnb_wgrad(C.byref(p), _lib.stream_ptr()))
    return dw


# ---- batch norm -------------------------------------------------------------------------------------------
_WS = {}


def reduce_ws(device, sid=0):
    """workspace of the deterministic reductions, one per (device, stream slot): zero-initialised once, the counters
    reset themselves; kernels sharing a workspace must be ordered on one stream"""
    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    key = (dev, sid)
    if key not in _WS:
        _WS[key] = torch.zeros(int(_lib.lib().hrnb_reduce_ws_floats()), dtype=torch.float32, device=dev)
    return _WS[key]


def bn_stats(c, sums, sid=0):
    _lib.check(_lib.lib().hrnb_bn_stats(c.ptr, c.ps, c.N, c.C, c.H, c.W, sums.data_ptr(),
                                        reduce_ws(sums.device, sid).data_ptr(), _lib.stream_ptr()))


def bn_params(c, sums, gamma, beta, out, res=None, relu=True, running_mean=None, running_var=None, out2=None):
    p = BnParams()
    p.c, p.c_ps, p.sums, p.gamma, p.beta = c.ptr, c.ps, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    p.res, p.res_ps = (res.ptr, res.ps) if res is not None else (None, 0)
    p.out, p.out_ps = out.ptr, out.ps
    p.running_mean = running_mean.data_ptr() if running_mean is not None else None
    p.running_var = running_var.data_ptr() if running_var is not None else None
    p.N, p.C, p.H, p.W, p.relu, p.eps, p.momentum = c.N, c.C, c.H, c.W, int(relu), EPS, MOMENTUM
    if out2 is not None:        # phase-split copy of `out` for a following stride-2 conv
        assert isinstance(out2, PhasePF8) and (out2.N, out2.C, out2.H, out2.W) == (c.N, c.C, c.H, c.W)
        p.out2, p.out2_ps, p.out2_phase_stride = out2.ptr, out2.ps, out2.phase_stride
    return p


def bn_apply(*a, **k):
    p = bn_params(*a, **k)
    _lib.check(_lib.lib().hrnb_bn_apply(C.byref(p), _lib.stream_ptr()))


def bn_bwd_params(dy, y, c, sums, gamma, dsums, dc, dgamma, dbeta, relu=True, dres=None, dres_mode=1, sid=0):
    p = BnBwdParams()
    p.dy, p.dy_ps = dy.ptr, dy.ps
    p.y, p.y_ps = (y.ptr, y.ps) if (y is not None and relu) else (None, 0)
    p.c, p.c_ps, p.sums, p.gamma, p.dsums = c.ptr, c.ps, sums.data_ptr(), gamma.data_ptr(), dsums.data_ptr()
    p.ws = reduce_ws(dsums.device, sid).data_ptr()
    p.dc, p.dc_ps = dc.ptr, dc.ps
    p.dres, p.dres_ps, p.dres_mode = (dres.ptr, dres.ps, dres_mode) if dres is not None else (None, 0, 0)
    p.dgamma = dgamma.data_ptr() if dgamma is not None else None
    p.dbeta = dbeta.data_ptr() if dbeta is not None else None
    p.N, p.C, p.H, p.W, p.relu, p.eps = c.N, c.C, c.H, c.W, int(relu), EPS
    return p


def bn_bwd(*a, **k):
    p = bn_bwd_params(*a, **k)
    _lib.check(_lib.lib().hrnb_bn_bwd_reduce(C.byref(p), _lib.stream_ptr()))
    _lib.check(_lib.lib().hrnb_bn_bwd_apply(C.byref(p), _lib.stream_ptr()))


# ---- elementwise backward ---------------------------------------------------------------------------------
def fuse_sum_bwd(dy, y, dsrc, shift, relu=True, mode=1):
    assert dsrc.H == dy.H >> shift and dsrc.W == dy.W >> shift and dsrc.C == dy.C
    _lib.check(_lib.lib().hrnb_fuse_sum_bwd(dy.ptr, dy.ps, y.ptr if y is not None else None, y.ps if y is not None else 0,
                                             dsrc.ptr, dsrc.ps, dy.N, dy.H, dy.W, dy.C, shift, int(relu), mode,
                                             _lib.stream_ptr()))


def fuse_sum_bwd_batch_args(dy, y, dsrcs, shifts, modes, relu=True):
    """-> argument tuple of hrnb_fuse_sum_bwd_batch (all sources of one fuse output in one launch); keep it alive"""
    n = len(dsrcs)
    assert 1 <= n <= 4 and all(d.H == dy.H >> s and d.W == dy.W >> s and d.C == dy.C for d, s in zip(dsrcs, shifts))
    ptrs = (C.c_void_p * n)(*[d.ptr for d in dsrcs])
    pss = (C.c_int64 * n)(*[d.ps for d in dsrcs])
    sh = (C.c_int32 * n)(*shifts)
    md = (C.c_int32 * n)(*modes)
    return (dy.ptr, dy.ps, y.ptr if y is not None else None, y.ps if y is not None else 0, n, ptrs, pss, sh, md,
            dy.N, dy.H, dy.W, dy.C, int(relu))


def fuse_sum_bwd_batch(dy, y, dsrcs, shifts, modes, relu=True):
    a = fuse_sum_bwd_batch_args(dy, y, dsrcs, shifts, modes, relu)
    _lib.check(_lib.lib().hrnb_fuse_sum_bwd_batch(*a, _lib.stream_ptr()))


def bilinear_up_bwd(d_dst, d_src, align_corners, mode=1):
    assert d_dst.C == d_src.C
    _lib.check(_lib.lib().hrnb_bilinear_up_bwd(d_dst.ptr, d_dst.ps, d_dst.N, d_dst.C, d_dst.H, d_dst.W, d_src.ptr, d_src.ps,
                                                d_src.H, d_src.W, int(bool(align_corners)), mode, _lib.stream_ptr()))


def phase_merge(src, dst, mode=1):
    assert isinstance(src, PhasePF8) and (src.N, src.C, src.H, src.W) == (dst.N, dst.C, dst.H, dst.W)
    _lib.check(_lib.lib().hrnb_phase_merge(src.ptr, src.ps, src.phase_stride, dst.ptr, dst.ps, dst.N, dst.C, dst.H, dst.W,
                                            mode, _lib.stream_ptr()))


def channel_sum(c, out, C_real, sid=0):
    _lib.check(_lib.lib().hrnb_channel_sum(c.ptr, c.ps, c.N, C_real, c.H, c.W, out.data_ptr(),
                                           reduce_ws(out.device, sid).data_ptr(), _lib.stream_ptr()))
