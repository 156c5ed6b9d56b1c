"""Thin Python wrappers over the C ABI ops (one call = one kernel launch on torch's current stream)."""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (ConvParams, FuseParams, HRNB_CONV_GATHER, HRNB_CONV_IN_PHASES, HRNB_CONV_OUT_NCHW,
                   HRNB_CONV_OUT_PHASES, HRNB_CONV_RELU, PackJob)
from .pf8 import PF8, PhasePF8

NUM_SMS = 148


def kc_candidates(cin):
    """even numbers of 8-channel planes per K chunk that divide cin/8, largest first"""
    planes = cin // 8
    if cin % 16:
        raise ValueError("cin must be a multiple of 16, got %d" % cin)
    return [kc for kc in (32, 16, 12, 8, 6, 4, 2) if kc <= planes and planes % kc == 0]


def pick_kc(cin):
    return kc_candidates(cin)[0]


def bn_candidates(cout):
    """legal N tiles: multiples of 16 that divide cout (<= 256); a single padded tile when cout % 16 != 0"""
    if cout % 16:
        if cout > 256:
            raise ValueError("cout=%d is neither a multiple of 16 nor <= 256" % cout)
        return [(cout + 15) // 16 * 16]
    return [bn for bn in range(min(cout, 256), 15, -16) if cout % bn == 0]


def pick_bn(cout):
    return bn_candidates(cout)[0]


def _smem_bytes(W, taps, mode, bn, mb, kc, custom=None, wres=False):
    """mode: 1 flat-shift, 2 gather (stride 2 or forced), 4 flat-shift over 4 input phases (stride 2);
    custom = (extra halo rows, sources) of a custom tap table"""
    gather = mode == 2
    if custom is not None:
        halo, nsrc = 128 * mb + custom[0], custom[1]
    elif mode == 4:
        halo, nsrc = 128 * mb + W + 2, 4
    else:
        halo, nsrc = 128 * mb + (2 * (W + 2) if (taps == 9 and not gather) else 0), 1
    a_stage = nsrc * kc * (128 * mb if gather else halo) * 16
    b_stage = taps * kc * bn * 16
    # wres (one K chunk, one N tile): the weights are loaded once per CTA into a single stage (conv_tc.cu)
    return 3456 + (5 if gather else 2) * a_stage + (1 if wres else 2) * b_stage


def _tile_model(P, W, cin, cout, taps, mode, has_res, bn, mb, kc, custom=None):
    """Rough cycle model of one conv launch for tile shape (bn, mb, kc): per-tile cost = max(tensor pipe incl. the
    shared-memory operand fetch and the per-chunk hand-off stalls, HBM/L2 bytes), times the tile rounds on 148 SMs."""
    mblocks = (P + 127) // 128
    tiles = (mblocks + mb - 1) // mb * (cout // bn if cout % 16 == 0 else 1)
    gather = mode == 2
    nchunks = (cin // 8) // kc
    # The model does NOT count the savings of resident weights / dual issue (conv_tc.cu: k.wres, k.dual): a picker that did
    # (HRNB_PICK_WRES=1) chose one-chunk tiles for the 64-channel layers and measured SLOWER in-trip on B200 (inference 21,197 vs
    # 22,032 images/s at batch 256, training 22.1 vs 21.8 ms/step) - fewer, fatter tiles win over resident weights there.
    wres = nchunks == 1 and cout == bn and os.environ.get("HRNB_PICK_WRES", "0") == "1"
    smem = _smem_bytes(W, taps, mode, bn, mb, kc, custom, wres)
    # opt-in (HRNB_SLAB=1, measured slower): 1x1 convs with several N tiles (the head conv) keep the whole weight slab of one
    # N tile resident, every CTA keeps to one N tile (conv_tc.cu: k.wres with nchunks > 1)
    n_tiles = cout // bn if cout % 16 == 0 else 1
    slab = False
    if not gather and taps == 1 and n_tiles > 1 and 1 < nchunks <= 8 and os.environ.get("HRNB_SLAB", "0") == "1":
        a_st = (_smem_bytes(W, taps, mode, bn, mb, kc, custom) - 3456 - 2 * taps * kc * bn * 16) // 2
        if 3456 + 2 * a_st + nchunks * taps * kc * bn * 16 <= 216 * 1024:
            slab, smem = True, 3456 + 2 * a_st + nchunks * taps * kc * bn * 16
    if smem > (216 if slab else 200) * 1024:
        return None
    ksteps = taps * cin // 16
    handoffs = nchunks * (taps if gather else 1)
    mma = ksteps * mb * max(bn / 2.0, (4096 + bn * 32) / 128.0) * 1.3 + 300.0 * handoffs + 400.0
    # dual issue (conv_tc.cu: k.dual): resident weights (one K chunk, one N tile) and room for three halo stages - the second
    # issuing warp hides the hand-off stalls and keeps the pipe fed
    a_stage = (smem - 3456 - (1 if wres else 2) * taps * kc * bn * 16) // (5 if gather else 2)
    if (not gather and wres and os.environ.get("HRNB_NO_DUAL", "0") != "1" and tiles >= 4 * NUM_SMS
            and 3456 + 3 * a_stage + taps * kc * bn * 16 <= 200 * 1024):
        mma = (mma - 300.0 * handoffs - 400.0) * 0.85 + 200.0
    if custom is not None:
        a_bytes = custom[1] * (128 * mb + custom[0]) * cin * 2
    elif mode == 4:
        a_bytes = 4 * (128 * mb + W + 2) * cin * 2
    elif gather:
        a_bytes = 128 * mb * taps * cin * 2 * 2        # 9 taps, half-used 32-byte sectors
    else:
        a_bytes = (128 * mb + (2 * (W + 2) if taps == 9 else 0)) * cin * 2
    io = a_bytes + (0 if (wres or slab) else bn * taps * cin * 2 * 0.5) + mb * 128 * bn * 2 * (2 if has_res else 1)
    cost = max(mma, io / 23.0)
    rounds = -(-tiles // NUM_SMS)        # one persistent CTA per SM
    return rounds * cost + 4500.0


def pick_tile(P, W, cin, cout, taps, mode, has_res, kc=None, custom=None):
    """-> (BN, MB, KC) minimising the launch-time model"""
    env_mb, env_bn, env_kc = os.environ.get("HRNB_MB"), os.environ.get("HRNB_BN"), os.environ.get("HRNB_KC")
    best = None
    kcs = [kc] if kc else kc_candidates(cin)
    if env_kc and int(env_kc) in kcs:
        kcs = [int(env_kc)]
    for bn in bn_candidates(cout):
        if env_bn and bn != int(env_bn) and int(env_bn) in bn_candidates(cout):
            continue
        for mb in (4, 2, 1):
            if mb * bn > 256 or (env_mb and mb != int(env_mb) and int(env_mb) * bn <= 256):
                continue
            for kcc in kcs:
                t = _tile_model(P, W, cin, cout, taps, mode, has_res, bn, mb, kcc, custom)
                if t is not None and (best is None or t < best[0]):
                    best = (t, bn, mb, kcc)
    if best is None:
        raise ValueError("no tile shape fits shared memory for cin=%d cout=%d" % (cin, cout))
    return best[1], best[2], best[3]


_TUNED = {}


def autotune_enabled():
    """opt-in (HRNB_AUTOTUNE=1): on B200 the measured-in-isolation winners (warm L2, back-to-back launches) were not
    faster in the real step than the cycle model's choice [29.6 vs 29.1 ms/step, round 1], so the model stays the default"""
    return os.environ.get("HRNB_AUTOTUNE", "0") == "1" and torch.cuda.is_available()


def _time_launch(fn, reps=3):
    """best-of-reps CUDA-event time (ms) of one launch on the current stream"""
    fn()
    best = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        t = a.elapsed_time(b)
        best = t if best is None else min(best, t)
    return best


def autotune_conv(layer, x, out, res, out2, max_candidates=8):
    """Measure the `max_candidates` tile shapes the cycle model ranks best for this conv (on the real buffers, with
    throw-away weight packs) and remember the fastest per shape.  Must not run under CUDA-graph capture: plans are
    built before they are captured.  Returns (BN, MB, KC) or None."""
    if torch.cuda.is_current_stream_capturing():
        return None
    H, W, P, mode, custom = layer._geometry(x)
    key = (layer.cin, layer.cout, layer.taps, layer.stride, mode, P, W, res is not None, custom, layer.out_nchw,
           isinstance(out, PhasePF8), out2 is not None, layer.flags)
    if key in _TUNED:
        return _TUNED[key]
    cands = []
    for bn in bn_candidates(layer.cout):
        for mb in (4, 2, 1):
            if mb * bn > 256:
                continue
            for kc in ([layer.fixed_kc] if layer.fixed_kc else kc_candidates(layer.cin)):
                t = _tile_model(P, W, layer.cin, layer.cout, layer.taps, mode, res is not None, bn, mb, kc, custom)
                if t is not None:
                    cands.append((t, bn, mb, kc))
    cands.sort()
    lib = _lib.lib()
    best = None
    if res is not None and not layer.out_nchw and res.ptr == out.ptr:
        out = PF8(out.N, out.C, out.H, out.W, device=out.buf.device)     # accumulate-in-place launch: tune into a scratch output
    for _, bn, mb, kc in cands[:max_candidates]:
        try:
            p = layer._build_params(x, out, res, out2, bn, mb, kc, temporary=True)
        except Exception:
            continue
        if p.MB != mb or lib.hrnb_conv_smem_bytes(C.byref(p)) < 0:
            continue
        ref = C.byref(p)
        ms = _time_launch(lambda: _lib.check(lib.hrnb_conv(ref, _lib.stream_ptr())))
        if best is None or ms < best[0]:
            best = (ms, bn, mb, kc)
    _TUNED[key] = (best[1], best[2], best[3]) if best else None
    return _TUNED[key]


def stats_eligible(p):
    """can this conv launch also produce the BatchNorm batch statistics of its output (include/hrnb.h: stats_sums)?"""
    widths = (16, 32, 64) if int(os.environ.get("HRNB_STATS_MAX", "64")) >= 64 else (16, 32)   # 8-epilogue-warp builds: <= 32
    return (p.BN == p.cout and p.cout in widths and not p.res and
            not (p.flags & (HRNB_CONV_GATHER | HRNB_CONV_OUT_NCHW | HRNB_CONV_RELU)))


def attach_stats(p, sums):
    """make conv launch `p` write (sum, sum of squares) per output channel into `sums` (fp32 [cout, 2]); the reduction
    workspace is private to the launch struct (zeroed once; the ticket counter resets itself).  False if not eligible."""
    if not stats_eligible(p):
        return False
    assert sums.dtype == torch.float32 and sums.numel() >= 2 * p.cout and sums.is_contiguous()
    ws = torch.zeros(int(_lib.lib().hrnb_conv_stats_ws_floats()), dtype=torch.float32, device=sums.device)
    p.stats_sums, p.stats_ws = sums.data_ptr(), ws.data_ptr()
    p._keep_stats = (sums, ws)
    return True


class ConvLayer:
    """conv (1x1 / 3x3 pad 1, stride 1 or 2) + folded BN (+ residual) (+ ReLU) on PF8 tensors.

    weight: OIHW fp32 (device); scale/shift: per-output-channel fp32 folded BN (None = 1 / 0).
    Weights are packed lazily per N-tile width (the tile shape depends on the batch/resolution it runs at).
    """

    def __init__(self, weight, scale=None, shift=None, stride=1, relu=False, out_nchw=False, kc=None, bn=None,
                 transpose=False, tap_ids=None, custom_taps=None, cin_pad=None, repacker=None):
        """General form (training path): `transpose` swaps the channel roles (data-gradient conv), `tap_ids` selects /
        reorders source taps, `custom_taps` = [(src, dpos)] gives the flat-shift geometry per logical tap
        (include/hrnb.h: ntap_custom), `cin_pad` zero-pads the logical input channels to a multiple of 16.  `weight`
        is NOT copied in the general form: packs are refreshed from it by `repacker.run()` after optimizer steps."""
        assert weight.is_cuda and weight.dtype == torch.float32
        cout, cin, kh, kw = weight.shape
        assert kh == kw and kh in (1, 3)
        self.src_cout, self.src_cin, self.src_taps = cout, cin, kh * kw
        self.transpose, self.tap_ids, self.custom_taps, self.repacker = transpose, tap_ids, custom_taps, repacker
        self.general = transpose or tap_ids is not None or custom_taps is not None or cin_pad is not None or repacker is not None
        if transpose:
            cout, cin = cin, cout
        if cin_pad is not None:
            cin = cin_pad
        ntap = len(tap_ids) if tap_ids is not None else kh * kw
        if custom_taps is not None:
            assert len(custom_taps) == ntap and stride == 1
        self.cout, self.cin, self.taps, self.stride = cout, cin, ntap, stride
        self.relu, self.out_nchw = relu, out_nchw
        self.fixed_kc = kc
        self.fixed_bn = bn
        self.w = weight if self.general else weight.contiguous()
        assert self.w.is_contiguous()
        self.sc = scale.contiguous().float() if scale is not None else None
        self.sh = shift.contiguous().float() if shift is not None else None
        self.packs = {}
        self.flags = (HRNB_CONV_RELU if relu else 0) | (HRNB_CONV_OUT_NCHW if out_nchw else 0) | \
                     (HRNB_CONV_GATHER if stride == 2 else 0)
        self.force_gather = False
        self.no_pdl = False

    def pack(self, bn, kc, temporary=False):
        """packed weights + bias for tile width bn / K chunk kc; temporary=True (autotuner) packs into throw-away buffers
        that are neither cached nor registered for re-packing"""
        if temporary:
            saved, saved_rp = self.packs, self.repacker
            self.packs, self.repacker = {}, None
            general = self.general
            self.general = True if saved_rp is not None else general
            try:
                return self.pack(bn, kc)
            finally:
                self.packs, self.repacker, self.general = saved, saved_rp, general
        if (bn, kc) not in self.packs:
            n_tiles = (self.cout + bn - 1) // bn
            dev = self.w.device
            wpk = torch.empty(n_tiles * bn * self.taps * self.cin, dtype=torch.bfloat16, device=dev)
            bias = torch.empty(n_tiles * bn, dtype=torch.float32, device=dev)
            if self.general:
                job = PackJob()
                job.w, job.scale = self.w.data_ptr(), (self.sc.data_ptr() if self.sc is not None else None)
                job.shift = self.sh.data_ptr() if self.sh is not None else None
                job.wpk_out, job.bias_out = wpk.data_ptr(), bias.data_ptr()
                job.cout, job.cin, job.taps_total = self.src_cout, self.src_cin, self.src_taps
                job.transpose, job.lcout, job.lcin, job.ntap = int(self.transpose), self.cout, self.cin, self.taps
                for t in range(self.taps):
                    job.tap_ids[t] = self.tap_ids[t] if self.tap_ids is not None else t
                job.KC, job.BN = kc, bn
                rp = self.repacker or Repacker(dev)
                rp.add(job, n_tiles * bn * self.taps * self.cin)
                if self.repacker is None:
                    rp.run()
                self.packs[(bn, kc)] = (wpk, bias)
                return self.packs[(bn, kc)]
            _lib.check(_lib.lib().hrnb_pack_conv_weights(
                self.w.data_ptr(), self.sc.data_ptr() if self.sc is not None else None,
                self.sh.data_ptr() if self.sh is not None else None, self.cout, self.cin, self.taps, kc, bn,
                wpk.data_ptr(), bias.data_ptr(), _lib.stream_ptr()))
            self.packs[(bn, kc)] = (wpk, bias)
        return self.packs[(bn, kc)]

    def _geometry(self, x):
        in_ph = isinstance(x, PhasePF8)
        H, W = x.H // self.stride, x.W // self.stride
        P = x.N * (H + 1) * (W + 1)
        if in_ph:
            assert self.stride == 2 and self.taps == 9 and not self.force_gather
            mode = 4
        else:
            mode = 2 if (self.stride == 2 or self.force_gather) else 1
        custom = None
        if self.custom_taps is not None:
            dps = [d for _, d in self.custom_taps]
            custom = (max(0, max(dps)) - min(0, min(dps)), max(s_ for s_, _ in self.custom_taps) + 1)
        return H, W, P, mode, custom

    def params(self, x, out, res=None, mb=None, bn=None, kc=None, out2=None, fuse=None, relu=None, up_shift=0, fuse_after=False):
        """launch parameters; the tile shape (BN, MB, KC) comes from the caller, else from the autotuner (measured on the
        device the first time a shape is seen, like the reference's cudnn.benchmark = True, tools/train.py:128), else from
        the cycle model"""
        if up_shift:
            return self._fuse_host_up(x, out, res, fuse, relu, up_shift)
        H, W, P, mode, custom = self._geometry(x)
        forced = mb is not None or bn is not None or kc is not None or self.fixed_bn is not None
        if not forced and autotune_enabled() and not fuse:
            tile = autotune_conv(self, x, out, res, out2)
            if tile is not None:
                return self._build_params(x, out, res, out2, *tile)
        tbn, tmb, tkc = pick_tile(P, W, self.cin, self.cout, self.taps, mode, res is not None, kc or self.fixed_kc, custom)
        bn = bn or self.fixed_bn or tbn
        if mb is None:
            mb = tmb if bn == tbn else 1
        while mb * bn > 256:
            mb //= 2
        kc = tkc
        if bn != tbn or mb != tmb:      # forced shape: take the largest K chunk that fits
            for cand in ([kc] if (kc and self.fixed_kc) else kc_candidates(self.cin)):
                if _smem_bytes(W, self.taps, mode, bn, mb, cand, custom) <= 200 * 1024:
                    kc = cand
                    break
        p = self._build_params(x, out, res, out2, bn, mb, kc)
        return self._attach_fuse(p, fuse, relu, fuse_after)

    @staticmethod
    def _attach_fuse(p, fuse, relu, fuse_after=False):
        """fuse: [(PF8 source, shift)] added in the epilogue with nearest up-sampling (include/hrnb.h: nfuse); relu: override
        of the layer's ReLU flag (the fuse-layer host conv applies the ReLU of the sum, pose_hrnet.py:266); fuse_after: the
        sources are added AFTER the unit's own ReLU and a second ReLU follows (HRNB_CONV_FUSE_AFTER_RELU)"""
        if fuse_after:
            assert fuse
            p.flags |= _lib.HRNB_CONV_FUSE_AFTER_RELU
        if fuse:
            assert len(fuse) <= 3
            p.nfuse = len(fuse)
            for i, (src, sh) in enumerate(fuse):
                assert src.C == p.cout and src.H == p.H >> sh and src.W == p.W >> sh, (src.C, src.H, src.W, sh)
                p.fuse_src[i], p.fuse_ps[i], p.fuse_shift[i] = src.ptr, src.ps, sh
            p._keep_fuse = [s_ for s_, _ in fuse]
        if relu is not None:
            p.flags = (p.flags | HRNB_CONV_RELU) if relu else (p.flags & ~HRNB_CONV_RELU)
        return p

    def _fuse_host_up(self, x, out, res, fuse, relu, up_shift):
        """1x1 conv evaluated on the grid of `out` with its input read through nearest up-sampling by 2^up_shift (gather
        path): the host convolution of the highest-resolution fuse output (include/hrnb.h: in_up_shift)"""
        assert self.taps == 1 and self.stride == 1 and x.H == out.H >> up_shift and x.W == out.W >> up_shift
        H, W = out.H, out.W
        P = x.N * (H + 1) * (W + 1)
        bn, mb, kc = pick_tile(P, W, self.cin, self.cout, 1, 2, res is not None, self.fixed_kc)
        wpk, bias = self.pack(bn, kc)
        p = ConvParams()
        p.inp, p.in_ps = x.ptr, x.ps
        p.wpk, p.bias = wpk.data_ptr(), bias.data_ptr()
        p.res, p.res_ps = (res.ptr, res.ps) if res is not None else (None, 0)
        p.out, p.out_ps = out.ptr, out.ps
        p.N, p.H, p.W, p.in_H, p.in_W = x.N, H, W, x.H, x.W
        p.cin, p.cout, p.taps, p.stride = self.cin, self.cout, 1, 1
        p.KC, p.BN, p.MB = kc, bn, mb
        p.flags = self.flags | HRNB_CONV_GATHER | (_lib.HRNB_CONV_NO_PDL if self.no_pdl else 0)
        p.in_up_shift = up_shift
        p._keep = (wpk, bias)
        while p.MB > 1 and _lib.lib().hrnb_conv_smem_bytes(C.byref(p)) < 0:
            p.MB //= 2
        return self._attach_fuse(p, fuse, relu)

    def _build_params(self, x, out, res, out2, bn, mb, kc, temporary=False):
        in_ph, out_ph = isinstance(x, PhasePF8), isinstance(out, PhasePF8)
        H, W = x.H // self.stride, x.W // self.stride
        lib = _lib.lib()
        wpk, bias = self.pack(bn, kc, temporary=temporary)
        p = ConvParams()
        p.inp, p.in_ps = x.ptr, x.ps
        p.in_phase_stride = x.phase_stride if in_ph else 0
        p.wpk, p.bias = wpk.data_ptr(), bias.data_ptr()
        p.res, p.res_ps = (res.ptr, res.ps) if res is not None else (None, 0)
        if self.out_nchw:
            p.out, p.out_ps = out.data_ptr(), 0
        else:
            assert out.H == H and out.W == W and out.C == self.cout and out.N == x.N
            p.out, p.out_ps = out.ptr, out.ps
        p.out_phase_stride = out.phase_stride if out_ph else 0
        if out2 is not None:      # second, phase-split copy of the output
            assert isinstance(out2, PhasePF8) and not out_ph and (out2.H, out2.W, out2.C) == (H, W, self.cout)
            p.out2, p.out2_ps, p.out2_phase_stride = out2.ptr, out2.ps, out2.phase_stride
        p.N, p.H, p.W, p.in_H, p.in_W = x.N, H, W, x.H, x.W
        p.cin, p.cout, p.taps, p.stride = self.cin, self.cout, self.taps, self.stride
        p.KC, p.BN, p.MB = kc, bn, mb
        flags = self.flags | (HRNB_CONV_GATHER if self.force_gather else 0)
        if in_ph:
            flags = (flags & ~HRNB_CONV_GATHER) | HRNB_CONV_IN_PHASES
        if out_ph:
            flags |= HRNB_CONV_OUT_PHASES
        if self.no_pdl:
            flags |= _lib.HRNB_CONV_NO_PDL
        if self.custom_taps is not None:
            p.ntap_custom = self.taps
            for t, (src, dpos) in enumerate(self.custom_taps):
                p.tap_src[t], p.tap_dpos[t] = src, dpos
        p.flags = flags
        p._keep = (wpk, bias)       # temporary packs live as long as the struct
        while p.MB > 1 and lib.hrnb_conv_smem_bytes(C.byref(p)) < 0:
            p.MB //= 2          # tile does not fit in shared memory at this MB
        return p

    def __call__(self, x, out, res=None, mb=None, bn=None, out2=None, stats=None, fuse=None, relu=None, up_shift=0, fuse_after=False):
        """stats: fp32 [cout, 2] tensor -> the launch also writes the BatchNorm batch statistics (sum, sum of squares per
        channel) of `out`; raises when the launch is not eligible (see attach_stats)"""
        assert x.C == self.cin, (x.C, self.cin)
        p = self.params(x, out, res, mb, bn, out2=out2, fuse=fuse, relu=relu, up_shift=up_shift, fuse_after=fuse_after)
        if stats is not None and not attach_stats(p, stats):
            raise ValueError("conv launch not eligible for fused BatchNorm statistics (BN=%d cout=%d flags=%d)" % (p.BN, p.cout, p.flags))
        _lib.check(_lib.lib().hrnb_conv(C.byref(p), _lib.stream_ptr()))
        return out


def grouped_conv_params(layers, x, outs, ress=None):
    """ONE launch for len(layers) <= 4 custom-tap convs over the same input `x` (include/hrnb.h: ngroup) - the per-phase
    data-gradient convs of a 3x3 stride-2 conv.  outs / ress: PF8 tensors (views) a constant stride apart."""
    n = len(layers)
    assert 2 <= n <= 4 and all(l.custom_taps is not None and l.cin == layers[0].cin and l.cout == layers[0].cout for l in layers)
    lead = max(range(n), key=lambda i: layers[i].taps)           # the widest conv decides the tile shape
    L = layers[lead]
    H, W, P, mode, custom = L._geometry(x)
    bn, mb, kc = pick_tile(P, W, L.cin, L.cout, L.taps, mode, ress is not None, L.fixed_kc, custom)
    p = layers[0]._build_params(x, outs[0], ress[0] if ress is not None else None, None, bn, mb, kc)
    p.ngroup = n
    keep = []
    for g, l in enumerate(layers):
        wpk, _ = l.pack(bn, kc)
        keep.append(wpk)
        p.grp_wpk[g], p.grp_ntap[g] = wpk.data_ptr(), l.taps
        for t, (src, dpos) in enumerate(l.custom_taps):
            assert src == 0 and t < 4
            p.grp_tap_dpos[g][t] = dpos
    stride = (outs[1].ptr - outs[0].ptr) // 2
    assert all((outs[g].ptr - outs[0].ptr) // 2 == g * stride and outs[g].ps == outs[0].ps for g in range(n))
    p.grp_out_stride = stride
    if ress is not None:
        rstride = (ress[1].ptr - ress[0].ptr) // 2
        assert all((ress[g].ptr - ress[0].ptr) // 2 == g * rstride and ress[g].ps == ress[0].ps for g in range(n))
        p.grp_res_stride = rstride
    p._keep_grp = keep
    while p.MB > 1 and _lib.lib().hrnb_conv_smem_bytes(C.byref(p)) < 0:
        p.MB //= 2
    return p


class Repacker:
    """Collects hrnb_pack_job records and (re)packs all of them with ONE launch of hrnb_pack_conv_weights_batch."""

    def __init__(self, device):
        self.device = device
        self.jobs, self.sizes = [], []
        self._dev = None

    def add(self, job, total_elems):
        self.jobs.append(job)
        self.sizes.append(int(total_elems))
        self._dev = None

    def _upload(self):
        import numpy as np
        block_job, b0 = [], 0
        arr = (PackJob * len(self.jobs))()
        for i, (job, n) in enumerate(zip(self.jobs, self.sizes)):
            nb = (n // job.ntap + 255) // 256       # one thread per (output channel, input channel) pair, all taps
            job.block0 = b0
            C.memmove(C.byref(arr, i * C.sizeof(PackJob)), C.byref(job), C.sizeof(PackJob))
            block_job.append(np.full(nb, i, dtype=np.int32))
            b0 += nb
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        self._dev = (torch.from_numpy(raw).to(self.device), torch.from_numpy(np.concatenate(block_job)).to(self.device), b0)

    def run(self):
        if not self.jobs:
            return
        if self._dev is None:
            self._upload()
        jobs, bmap, nb = self._dev
        _lib.check(_lib.lib().hrnb_pack_conv_weights_batch(jobs.data_ptr(), bmap.data_ptr(), nb, _lib.stream_ptr()))


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


def phase_split(src, dst):
    """PF8 [N,C,H,W] -> PhasePF8 (4 half-resolution tensors) for a following 3x3 stride-2 conv."""
    assert isinstance(dst, PhasePF8) and (src.N, src.C, src.H, src.W) == (dst.N, dst.C, dst.H, dst.W)
    _lib.check(_lib.lib().hrnb_phase_split(src.ptr, src.ps, src.N, src.C, src.H, src.W, dst.ptr, dst.ps,
                                           dst.phase_stride, _lib.stream_ptr()))
    return dst


def stem_im2col(x, out):
    """x: [N,3,H,W] fp32 NCHW -> PF8 32 channels (27 taps + 5 zeros) at H/2 x W/2."""
    N, c, H, W = x.shape
    assert c == 3 and out.C == 32 and out.H == H // 2 and out.W == W // 2
    _lib.check(_lib.lib().hrnb_stem_im2col(x.data_ptr(), out.ptr, out.ps, N, H, W, _lib.stream_ptr()))
    return out
