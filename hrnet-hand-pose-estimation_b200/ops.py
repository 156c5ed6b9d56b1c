"""Thin Python wrappers over the C ABI ops (one call = one kernel launch on torch's current stream)."""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import ConvParams, FuseParams, HRNB_CONV_GATHER, HRNB_CONV_OUT_NCHW, HRNB_CONV_RELU
from .pf8 import PF8

NUM_SMS = 148


def pick_kc(cin):
    planes = cin // 8
    for kc in (8, 6, 4, 2):
        if planes % kc == 0:
            return kc
    raise ValueError("cin must be a multiple of 16, got %d" % cin)


def pick_bn(cout):
    if cout % 16:
        return (cout + 15) // 16 * 16 if cout < 256 else None
    if cout <= 256:
        return cout
    for bn in (256, 240, 224, 208, 192, 176, 160, 144, 128):
        if cout % bn == 0:
            return bn
    raise ValueError("no N tile for cout=%d" % cout)


def pick_mb(P, BN, taps, n_tiles):
    """M blocks per CTA: amortise the 3x3 halo / weight tiles while keeping >= ~2 waves of CTAs."""
    env = os.environ.get("HRNB_MB")
    if env:
        mb = int(env)
        while mb * BN > 512:
            mb //= 2
        return max(mb, 1)
    mblocks = (P + 127) // 128
    best = 1
    for mb in (2, 4):
        if mb * BN > 512:
            break
        if (mblocks + mb - 1) // mb * n_tiles >= 2 * NUM_SMS:
            best = mb
    return best


class ConvLayer:
    """conv (1x1 / 3x3 pad 1, stride 1 or 2) + folded BN (+ residual) (+ ReLU) on PF8 tensors.

    weight: OIHW fp32 (device); scale/shift: per-output-channel fp32 folded BN (None = 1 / 0).
    """

    def __init__(self, weight, scale=None, shift=None, stride=1, relu=False, out_nchw=False, kc=None, bn=None):
        assert weight.is_cuda and weight.dtype == torch.float32
        cout, cin, kh, kw = weight.shape
        assert kh == kw and kh in (1, 3)
        self.cout, self.cin, self.taps, self.stride = cout, cin, kh * kw, stride
        self.relu, self.out_nchw = relu, out_nchw
        self.KC = kc or pick_kc(cin)
        self.BN = bn or pick_bn(cout)
        self.n_tiles = (cout + self.BN - 1) // self.BN
        dev = weight.device
        self.wpk = torch.empty(self.n_tiles * self.BN * self.taps * cin, dtype=torch.bfloat16, device=dev)
        self.bias = torch.empty(self.n_tiles * self.BN, dtype=torch.float32, device=dev)
        w = weight.contiguous()
        sc = scale.contiguous().float() if scale is not None else None
        sh = shift.contiguous().float() if shift is not None else None
        _lib.check(_lib.lib().hrnb_pack_conv_weights(
            w.data_ptr(), sc.data_ptr() if sc is not None else None, sh.data_ptr() if sh is not None else None,
            cout, cin, self.taps, self.KC, self.BN, self.wpk.data_ptr(), self.bias.data_ptr(), _lib.stream_ptr()))
        self.flags = (HRNB_CONV_RELU if relu else 0) | (HRNB_CONV_OUT_NCHW if out_nchw else 0) | \
                     (HRNB_CONV_GATHER if stride == 2 else 0)
        self.force_gather = False

    def params(self, x, out, res=None, mb=None):
        H, W = x.H // self.stride, x.W // self.stride
        p = ConvParams()
        p.inp, p.in_ps = x.ptr, x.ps
        p.wpk, p.bias = self.wpk.data_ptr(), self.bias.data_ptr()
        p.res, p.res_ps = (res.ptr, res.ps) if res is not None else (None, 0)
        if self.out_nchw:
            p.out, p.out_ps = out.data_ptr(), 0
        else:
            assert out.H == H and out.W == W and out.C == self.cout and out.N == x.N
            p.out, p.out_ps = out.ptr, out.ps
        p.N, p.H, p.W, p.in_H, p.in_W = x.N, H, W, x.H, x.W
        p.cin, p.cout, p.taps, p.stride = self.cin, self.cout, self.taps, self.stride
        p.KC, p.BN = self.KC, self.BN
        P = x.N * (H + 1) * (W + 1)
        p.MB = mb or pick_mb(P, self.BN, self.taps, self.n_tiles)
        p.flags = self.flags | (HRNB_CONV_GATHER if self.force_gather else 0)
        lib = _lib.lib()
        while p.MB > 1 and lib.hrnb_conv_smem_bytes(C.byref(p)) < 0:
            p.MB //= 2          # tile does not fit in shared memory at this MB
        return p

    def __call__(self, x, out, res=None, mb=None):
        assert x.C == self.cin, (x.C, self.cin)
        p = self.params(x, out, res, mb)
        _lib.check(_lib.lib().hrnb_conv(C.byref(p), _lib.stream_ptr()))
        return out


def fuse_sum(srcs, shifts, out, relu=True):
    p = FuseParams()
    for i, (s, sh) in enumerate(zip(srcs, shifts)):
        p.src[i], p.src_ps[i], p.shift[i] = s.ptr, s.ps, sh
        assert s.C == out.C and s.H == out.H >> sh and s.W == out.W >> sh
    p.nsrc = len(srcs)
    p.out, p.out_ps = out.ptr, out.ps
    p.N, p.H, p.W, p.C, p.relu = out.N, out.H, out.W, out.C, int(relu)
    _lib.check(_lib.lib().hrnb_fuse_sum(C.byref(p), _lib.stream_ptr()))
    return out


def bilinear_up(src, dst, align_corners):
    assert src.C == dst.C and src.N == dst.N
    _lib.check(_lib.lib().hrnb_bilinear_up(src.ptr, src.ps, src.N, src.C, src.H, src.W, dst.ptr, dst.ps,
                                           dst.H, dst.W, int(bool(align_corners)), _lib.stream_ptr()))
    return dst


def stem_conv1(x, w27, bias, out):
    """x: [N,3,H,W] fp32 NCHW; w27: [64,27] fp32 (BN scale folded); out: PF8 64 ch at H/2 x W/2."""
    N, c, H, W = x.shape
    assert c == 3 and out.C == 64 and out.H == H // 2 and out.W == W // 2
    _lib.check(_lib.lib().hrnb_stem_conv1(x.data_ptr(), w27.data_ptr(), bias.data_ptr(), out.ptr, out.ps,
                                          N, H, W, _lib.stream_ptr()))
    return out
