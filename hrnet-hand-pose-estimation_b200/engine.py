"""Inference engine: turns a reference-layout state_dict into packed bf16 weights + a launch plan.

Host side of the hot path.  Everything numerical happens in libhrnb.so; this file only decides which
kernel runs on which buffer and stream:

  * BN (eval) is folded: scale into the bf16 weights, shift into the fp32 epilogue bias
    (BatchNorm2d eps 1e-5: lib/models/pose_hrnet.py:36 et al.).
  * every conv is one launch with bias / residual / ReLU fused (BasicBlock :43-59, Bottleneck :80-100).
  * a fuse layer = low-resolution 1x1 convs + stride-2 chains + ONE sum kernel per output that applies the
    nearest up-sampling on the fly and the ReLU (HighResolutionModule.forward :247-266) - up-sampled
    tensors are never materialised.
  * the head writes branch 0 straight into the concat buffer and bilinearly up-samples branches 1-3 into
    it (:560-565 / pose_hrnet_softmax.py:499-504).
  * branches of a module run on separate CUDA streams; the whole plan is replayed as one CUDA graph.
"""
import ctypes as C
import os
import threading

import torch

from . import _lib, arch as A
from .ops import ConvLayer, PF8, PhasePF8
from ._lib import FuseParams

EPS = 1e-5
_BUILD_LOCK = threading.RLock()    # engine / plan construction and graph capture: one thread at a time (nn.DataParallel replicas)


def _fold(sd, bn_key, conv_bias=None, device=None):
    """eval-mode BatchNorm as per-channel (scale, shift), computed on the HOST in fp32 (a few hundred floats per layer) and
    uploaded: the device only ever runs the library's own kernels while an engine is built (the pack kernel folds `scale`
    into the bf16 weights, `shift` becomes the epilogue bias)"""
    g, b = sd[bn_key + ".weight"].float().cpu(), sd[bn_key + ".bias"].float().cpu()
    m, v = sd[bn_key + ".running_mean"].float().cpu(), sd[bn_key + ".running_var"].float().cpu()
    scale = g / torch.sqrt(v + EPS)
    shift = b - m * scale
    if conv_bias is not None:
        shift = shift + conv_bias.float().cpu() * scale
    return scale.to(device), shift.to(device)


class _Step:
    __slots__ = ("kind", "sid", "fn", "other", "name")

    def __init__(self, kind, sid, fn=None, other=None, name=""):
        self.kind, self.sid, self.fn, self.other, self.name = kind, sid, fn, other, name


class Plan:
    """Buffers + ordered launch list for one (batch, H, W)."""

    def __init__(self, engine, B, H, W):
        self.engine, self.B, self.H, self.W = engine, B, H, W
        self.steps = []
        self.keep = []          # keeps ctypes structs / tensors alive
        self.graph = None
        self.n_launch = 0
        dev = engine.device
        self.x = torch.zeros((B, 3, H, W), dtype=torch.float32, device=dev)
        self.out = {}
        self._build()

    # ---- step recording ---------------------------------------------------------------------------
    def _op(self, sid, fn, name):
        if self.engine.single_stream:
            sid = 0
        self.steps.append(_Step("op", sid, fn, name=name))
        self.n_launch += 1

    def _wait(self, sid, other):
        if sid != other and not self.engine.single_stream:
            self.steps.append(_Step("wait", sid, other=other))

    def _buf(self, C_, H, W):
        t = PF8(self.B, C_, H, W, device=self.engine.device)
        self.keep.append(t)
        return t

    def _phases(self, C_, H, W):
        t = PhasePF8(self.B, C_, H, W, device=self.engine.device)
        self.keep.append(t)
        return t

    def _split(self, sid, src, name=""):
        """PF8 -> PhasePF8 copy feeding the 3x3 stride-2 convs that read `src` (one split serves all of them)."""
        dst = self._phases(src.C, src.H, src.W)
        lib = _lib.lib()

        def fn(lib=lib, s=src, d=dst):
            _lib.check(lib.hrnb_phase_split(s.ptr, s.ps, s.N, s.C, s.H, s.W, d.ptr, d.ps, d.phase_stride,
                                            _lib.stream_ptr()))
        self._op(sid, fn, name)
        return dst

    def _conv(self, sid, layer, x, out, res=None, name="", out2=None, fuse=None, relu=None, up_shift=0, fuse_after=False):
        layer.no_pdl = not self.engine.pdl
        p = layer.params(x, out, res, out2=out2, fuse=fuse, relu=relu, up_shift=up_shift, fuse_after=fuse_after)
        self.keep.append(p)
        lib = _lib.lib()
        ref = C.byref(p)

        def fn(lib=lib, ref=ref):
            _lib.check(lib.hrnb_conv(ref, _lib.stream_ptr()))
        self._op(sid, fn, name)
        return out

    def _fuse(self, sid, srcs, shifts, out, name=""):
        p = FuseParams()
        for i, (s, sh) in enumerate(zip(srcs, shifts)):
            p.src[i], p.src_ps[i], p.shift[i] = s.ptr, s.ps, sh
        p.nsrc = len(srcs)
        p.out, p.out_ps = out.ptr, out.ps
        p.N, p.H, p.W, p.C, p.relu = out.N, out.H, out.W, out.C, 1
        self.keep.append(p)
        lib = _lib.lib()
        ref = C.byref(p)

        def fn(lib=lib, ref=ref):
            _lib.check(lib.hrnb_fuse_sum(ref, _lib.stream_ptr()))
        self._op(sid, fn, name)
        return out

    def _down_chain(self, pre, i, j, t, ch, res_hw, last_kw=None):
        """fuse_layers.i.j for j < i: (i - j) stride-2 3x3 convs from branch j's phase-split output `t`; intermediate outputs
        are written as phases by the conv epilogue itself.  last_kw: extra arguments of the LAST conv (fuse-sum host)"""
        L, e = self.engine.layers, self.engine
        for k in range(i - j):
            fp = "%s.fuse_layers.%d.%d.%d.0" % (pre, i, j, k)
            last = k == i - j - 1
            co = ch[i] if last else ch[j]
            if last and last_kw is not None:
                return self._conv(i, L[fp], t, name=fp, **last_kw)
            dst = self._buf(co, *res_hw[j + k + 1]) if last else self._phases(co, *res_hw[j + k + 1])
            t = self._conv(i, L[fp], t, dst, name=fp)
        return t

    def _up_tree(self, pre, i, nb, xs, ch, res_hw, sids):
        """The up-path terms of fuse output i, sum_{j>i} up_{2^(j-i)}(f_ij(x_j)), as ONE tensor on branch i+1's grid: the 1x1 convs
        run from the lowest resolution upwards and each adds the 2x up-sampled result of the previous one in its epilogue
        (nearest up-sampling by powers of two composes exactly), so the full-resolution host conv reads one extra source
        instead of up to three - its epilogue is issue-bound, every source costs ~65 instructions per item.  sids[j]: stream of
        the conv from branch j.  None when i is the last branch."""
        L, z = self.engine.layers, None
        for j in range(nb - 1, i, -1):
            fp = "%s.fuse_layers.%d.%d.0" % (pre, i, j)
            if z is not None:
                self._wait(sids[j], sids[j + 1])
            z = self._conv(sids[j], L[fp], xs[j], self._buf(ch[i], *res_hw[j]), name=fp + ("+up" if z is not None else ""),
                           fuse=[(z, 1)] if z is not None else None)
        return z

    def _fuse_in_epilogue(self, pre, i, nb, xs, split, dst, ch, res_hw):
        """Fuse output i = ReLU(sum_j f_ij(x_j)) (lib/models/pose_hrnet.py:199-207,257-266) WITHOUT a separate sum pass: one
        of the output's own convolutions is the host of the sum - its epilogue adds the identity branch (residual input) and
        the other contributions (nearest up-sampled on the fly) and applies the ReLU.  Host: the single-step stride-2 conv
        from branch i-1 (i >= 1); for i = 0, which has no stride-2 chain, the 1x1 conv from branch 1 evaluated on the 64x64
        grid with its input up-sampled (1x1 conv and nearest up-sampling commute)."""
        L = self.engine.layers
        fuse = []
        if self.engine.fuse_tree and i > 0:
            z = self._up_tree(pre, i, nb, xs, ch, res_hw, [i] * nb)
            if z is not None:
                fuse.append((z, 1))
        for j in range(nb):
            if j > i and not (i == 0 and j == 1):
                if self.engine.fuse_tree and i > 0:
                    continue
                fp = "%s.fuse_layers.%d.%d.0" % (pre, i, j)
                fuse.append((self._conv(i, L[fp], xs[j], self._buf(ch[i], *res_hw[j]), name=fp), j - i))
            elif j < i - 1:
                fuse.append((self._down_chain(pre, i, j, split[j], ch, res_hw), 0))
        if i == 0:
            fp = "%s.fuse_layers.0.1.0" % pre
            return self._conv(0, L[fp], xs[1], dst, res=xs[0], name=fp + "+fuse", fuse=fuse, relu=True, up_shift=1)
        return self._down_chain(pre, i, i - 1, split[i - 1], ch, res_hw,
                                last_kw=dict(out=dst, res=xs[i], fuse=fuse, relu=True))

    # ---- the network -------------------------------------------------------------------------------
    def _build(self):
        e, L, B = self.engine, self.engine.layers, self.B
        arch, lib = e.arch, _lib.lib()
        ch = arch.channels
        H2, W2, H4, W4 = self.H // 2, self.W // 2, self.H // 4, self.W // 4
        if self.H % 32 or self.W % 32:
            raise ValueError("input H and W must be multiples of 32 (four resolutions, each halving)")

        # stem: conv1 as im2col (27 -> 32 channel slab) + 1x1 conv on the tensor pipe, written directly as the 4
        # phases conv2 (3x3 stride 2) reads; conv2 then runs on the flat-shift path
        cols = self._buf(32, H2, W2)
        x = self.x

        def stem(lib=lib, x=x, cols=cols, B=B, H=self.H, W=self.W):
            _lib.check(lib.hrnb_stem_im2col(x.data_ptr(), cols.ptr, cols.ps, B, H, W, _lib.stream_ptr()))
        self._op(0, stem, "conv1.im2col")
        # uint8 NHWC input with ToTensor + Normalize folded into the im2col (lib/dataset/transforms/build.py:82-85): SURVEY §8 f3
        self.x_u8 = None
        self.norm = None

        def stem_u8(lib=lib, cols=cols, B=B, H=self.H, W=self.W):
            mean, std = self.norm
            _lib.check(lib.hrnb_stem_im2col_u8(self.x_u8.data_ptr(), mean, std, cols.ptr, cols.ps, B, H, W, _lib.stream_ptr()))
        self.stem_u8 = stem_u8
        t1 = self._conv(0, L["conv1"], cols, self._phases(64, H2, W2), name="conv1")
        cur = self._conv(0, L["conv2"], t1, self._buf(64, H4, W4), name="conv2")

        # layer1: 4 bottlenecks
        for b in range(4):
            pre = "layer1.%d" % b
            c1 = self._conv(0, L[pre + ".conv1"], cur, self._buf(64, H4, W4), name=pre + ".conv1")
            c2 = self._conv(0, L[pre + ".conv2"], c1, self._buf(64, H4, W4), name=pre + ".conv2")
            res = cur
            if b == 0:
                res = self._conv(0, L[pre + ".downsample.0"], cur, self._buf(256, H4, W4), name=pre + ".downsample")
            cur_ph = self._phases(256, H4, W4) if b == 3 else None     # phase copy for the stride-2 transition conv
            cur = self._conv(0, L[pre + ".conv3"], c2, self._buf(256, H4, W4), res=res, name=pre + ".conv3", out2=cur_ph)

        # transition1
        res_hw = [(H4 >> i, W4 >> i) for i in range(4)]
        xs = [self._conv(0, L["transition1.0.0"], cur, self._buf(ch[0], *res_hw[0]), name="transition1.0")]
        self._wait(1, 0)
        xs.append(self._conv(1, L["transition1.1.0.0"], cur_ph, self._buf(ch[1], *res_hw[1]), name="transition1.1"))

        cat = None
        stage3_b0 = None
        for s, nmod in zip((2, 3, 4), arch.modules):
            nb = s
            if s > 2:
                key = "transition%d.%d.0.0" % (s - 1, nb - 1)
                tph = self._split(nb - 2, xs[-1], key + ".split")
                self._wait(nb - 1, nb - 2)
                xs.append(self._conv(nb - 1, L[key], tph, self._buf(ch[nb - 1], *res_hw[nb - 1]), name=key))
            for m in range(nmod):
                pre = "stage%d.%d" % (s, m)
                last_module = (s == 4 and m == nmod - 1)
                # branches: 4 BasicBlocks each, branch i on stream i
                split = {}   # branch outputs that feed stride-2 chains also exist as phases (written by the last conv)
                host0 = None  # fuse output 0 hosted by the LAST conv of branch 0 (deferred until the other branches are done)
                for i in range(nb):
                    for b in range(arch.blocks):
                        bp = "%s.branches.%d.%d" % (pre, i, b)
                        y = self._conv(i, L[bp + ".conv1"], xs[i], self._buf(ch[i], *res_hw[i]), name=bp + ".conv1")
                        ph = None
                        if b == arch.blocks - 1 and i < nb - 1:
                            ph = split[i] = self._phases(ch[i], *res_hw[i])
                        if b == arch.blocks - 1 and i == 0 and e.fuse_epilogue and e.fuse_host0 == "conv2":
                            host0 = (L[bp + ".conv2"], y, xs[0], ph, bp + ".conv2+fuse")
                            continue
                        xs[i] = self._conv(i, L[bp + ".conv2"], y, self._buf(ch[i], *res_hw[i]), res=xs[i],
                                           name=bp + ".conv2", out2=ph)
                if last_module:
                    cat = self._buf(arch.head_channels, *res_hw[0])
                out0 = None
                if host0 is not None:
                    # Output 0 = ReLU(x0 + sum_j up(f_0j(x_j))) with x0 = ReLU(conv2(...) + residual): the low-resolution 1x1 convs
                    # f_0j run first (their inputs are ready long before branch 0's eighth conv), then branch 0's last conv adds
                    # them AFTER its own ReLU (HRNB_CONV_FUSE_AFTER_RELU) - x0 itself only leaves the SM as the phase-split
                    # copy the stride-2 chains read; the full-resolution sum never makes an extra trip through HBM.
                    if e.fuse_tree:
                        fuse0 = [(self._up_tree(pre, 0, nb, xs, ch, res_hw, list(range(nb))), 1)]
                        self._wait(0, 1)
                    else:
                        fuse0 = []
                        for j in range(1, nb):
                            fp = "%s.fuse_layers.0.%d.0" % (pre, j)
                            fuse0.append((self._conv(j, L[fp], xs[j], self._buf(ch[0], *res_hw[j]), name=fp), j))
                        for j in range(1, nb):
                            self._wait(0, j)
                    layer, y, res0, ph, name = host0
                    out0 = cat.view_planes(0, ch[0] // 8) if last_module else self._buf(ch[0], *res_hw[0])
                    self._conv(0, layer, y, out0, res=res0, name=name, out2=ph, fuse=fuse0, relu=True, fuse_after=True)
                    xs[0] = None
                # every fuse output needs every branch
                for i in range(nb):
                    for j in range(nb):
                        self._wait(i, j)
                outs = []
                for i in range(nb):
                    if i == 0 and out0 is not None:
                        outs.append(out0)
                        continue
                    if last_module and i == 0:
                        dst = cat.view_planes(0, ch[0] // 8)
                    else:
                        dst = self._buf(ch[i], *res_hw[i])
                    if e.fuse_epilogue:
                        outs.append(self._fuse_in_epilogue(pre, i, nb, xs, split, dst, ch, res_hw))
                        continue
                    srcs, shifts = [], []
                    for j in range(nb):
                        if j == i:
                            srcs.append(xs[j]); shifts.append(0)
                        elif j > i:
                            fp = "%s.fuse_layers.%d.%d.0" % (pre, i, j)
                            z = self._conv(i, L[fp], xs[j], self._buf(ch[i], *res_hw[j]), name=fp)
                            srcs.append(z); shifts.append(j - i)
                        else:
                            srcs.append(self._down_chain(pre, i, j, split[j], ch, res_hw)); shifts.append(0)
                    outs.append(self._fuse(i, srcs, shifts, dst, name="%s.fuse.%d" % (pre, i)))
                # the next module's branch j reads outs[j], produced on stream j: no extra sync needed;
                # the fuse kernels of other streams still read the old xs, which are distinct buffers.
                xs = outs
            if s == 3:
                stage3_b0 = xs[0]

        # head: bilinear up-sample branches 1..3 into the concat buffer, each on its own stream
        align = 1 if e.variant == "softmax" else 0
        plane0 = ch[0] // 8
        for i in range(1, 4):
            dst = cat.view_planes(plane0, ch[i] // 8)
            plane0 += ch[i] // 8
            src = xs[i]

            def up(lib=lib, src=src, dst=dst, align=align):
                _lib.check(lib.hrnb_bilinear_up(src.ptr, src.ps, src.N, src.C, src.H, src.W, dst.ptr, dst.ps, dst.H,
                                                dst.W, align, _lib.stream_ptr()))
            self._op(i, up, "head.bilinear.%d" % i)
        for i in range(1, 4):
            self._wait(0, i)
        hid = self._conv(0, L["last_layer.0"], cat, self._buf(arch.head_channels, *res_hw[0]), name="last_layer.0")
        J = arch.num_joints
        logits = torch.empty((B, J, H4, W4), dtype=torch.float32, device=e.device)
        self._conv(0, L["last_layer.3"], hid, logits, name="last_layer.3")
        self.out["logits"] = logits
        self.cat, self.stage3_b0 = cat, stage3_b0

        if e.variant == "softmax":
            heat = torch.empty_like(logits)
            coords = torch.empty((B, J, 2), dtype=torch.float32, device=e.device)
            temp = e.temp

            def sm(lib=lib, logits=logits, temp=temp, heat=heat, coords=coords, BJ=B * J, h=H4, w=W4):
                _lib.check(lib.hrnb_softmax_softargmax(logits.data_ptr(), temp.data_ptr(), BJ, h, w, heat.data_ptr(),
                                                       coords.data_ptr(), _lib.stream_ptr()))
            self._op(0, sm, "softmax_softargmax")
            self.out["heatmap"], self.out["coords"] = heat, coords
        else:
            preds = torch.empty((B, J, 2), dtype=torch.float32, device=e.device)
            maxvals = torch.empty((B, J, 1), dtype=torch.float32, device=e.device)

            def am(lib=lib, logits=logits, preds=preds, maxvals=maxvals, BJ=B * J, h=H4, w=W4):
                _lib.check(lib.hrnb_decode_argmax(logits.data_ptr(), BJ, h, w, 0, 1, preds.data_ptr(),
                                                  maxvals.data_ptr(), None, _lib.stream_ptr()))
            self._op(0, am, "decode_argmax")
            self.out["preds"], self.out["maxvals"] = preds, maxvals

        # optional feature output (NCHW fp32, as the reference returns it)
        feat_src = cat if e.variant == "softmax" else stage3_b0
        feat = torch.empty((B, feat_src.C, feat_src.H, feat_src.W), dtype=torch.float32, device=e.device)

        def tofeat(lib=lib, s=feat_src, feat=feat):
            _lib.check(lib.hrnb_pf8_to_nchw_f32(s.ptr, s.ps, s.N, s.C, s.H, s.W, feat.data_ptr(), _lib.stream_ptr()))
        self.feat_step = _Step("op", 0, tofeat, name="features_to_nchw")
        self.out["features"] = feat

    # ---- execution ----------------------------------------------------------------------------------
    def _run_steps(self, want_features, u8=False):
        main = torch.cuda.current_stream()
        side = self.engine.side_streams
        streams = [main] + side
        for sd_ in side:
            sd_.wait_stream(main)
        for st in self.steps:
            if st.kind == "op":
                if st.sid == 0:
                    (self.stem_u8 if (u8 and st.name == "conv1.im2col") else st.fn)()
                else:
                    with torch.cuda.stream(streams[st.sid]):
                        st.fn()
            else:
                streams[st.sid].wait_stream(streams[st.other])
        for sd_ in side:
            main.wait_stream(sd_)
        if want_features:
            self.feat_step.fn()

    def launches(self, want_features):
        return self.n_launch + (1 if want_features else 0)

    def run(self, want_features=True, use_graph=True, u8=False):
        if not use_graph:
            self._run_steps(want_features, u8)
            return self.out
        key = (bool(want_features), bool(u8))
        if self.graph is None:
            self.graph = {}
        if key not in self.graph:
            # warm-up outside capture (sets function attributes, loads modules), then capture.  nn.DataParallel calls forward from
            # one thread per GPU: building (device synchronisation, allocation) and capturing are serialised across threads -
            # a cudaDeviceSynchronize of one replica thread while another one captured failed with
            # cudaErrorStreamCaptureUnsupported about one run in three, thread-local capture mode notwithstanding
            with _BUILD_LOCK:
                self._run_steps(want_features, u8)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._run_steps(want_features, u8)
                self.graph[key] = g
        self.graph[key].replay()
        return self.out


class HRNetEngine:
    def __init__(self, sd, arch, variant, device):
        self.arch, self.variant, self.device = arch, variant, torch.device(device)
        self.layers = {}
        self.plans = {}
        self.use_graph = os.environ.get("HRNB_NO_GRAPH", "0") != "1"
        self.single_stream = os.environ.get("HRNB_SINGLE_STREAM", "0") == "1"
        # programmatic dependent launch of the conv kernels: default on since the round-2 soak (train.py); neutral for the
        # graph-replayed inference plan (13.12 vs 13.13 ms at batch 256), kept on so both engines run one launch mode.  HRNB_PDL=0: off
        self.pdl = os.environ.get("HRNB_PDL", "1") != "0"
        # fuse-layer sums inside the epilogue of one of the output's own convs (no separate fuse_sum pass, no second trip of the
        # summed tensor through HBM); HRNB_FUSE_EPILOGUE=0 restores the stand-alone sum kernel
        self.fuse_epilogue = os.environ.get("HRNB_FUSE_EPILOGUE", "1") != "0"
        # host of fuse output 0: "conv2" = the last conv of branch 0 (default), "gather" = the 1x1 conv from branch 1 evaluated
        # on the up-sampled grid (round-2 first form, 88 us instead of ~10 us extra at batch 256)
        self.fuse_host0 = os.environ.get("HRNB_FUSE_HOST0", "conv2")
        # HRNB_FUSE_TREE=1: up-path terms of a fuse output pre-summed on the low-resolution grids (Plan._up_tree) instead of
        # every 1x1 conv output being a separate source of the host conv.  Opt-in: +0.5 % at batch 256, -0.8 % at batch 64,
        # -3 % at batch 8 in-trip (the chain of small convs is serial, the separate convs run side by side)
        self.fuse_tree = os.environ.get("HRNB_FUSE_TREE", "0") == "1"
        with torch.cuda.device(self.device), _BUILD_LOCK:
            _lib.hang_init()
            self.side_streams = [torch.cuda.Stream(device=self.device) for _ in range(3)]
            self._pack(sd)

    def _pack(self, sd):
        specs = A.layer_specs(self.arch)
        bn_after = {}
        for a, b in zip(specs[:-1], specs[1:]):
            if isinstance(a, A.Conv) and isinstance(b, A.BN):
                bn_after[a.key] = b.key
        relu_less = set()   # convs whose BN output feeds a sum (no ReLU): block conv2/conv3, downsample, fuse finals
        for sp in specs:
            if not isinstance(sp, A.Conv):
                continue
            k = sp.key
            leaf = k.rsplit(".", 1)[-1]
            if ".fuse_layers." in k:
                parts = k.split(".")          # stageS.M.fuse_layers.I.J[.K].0
                i, j = int(parts[3]), int(parts[4])
                if j > i or int(parts[5]) == i - j - 1:
                    relu_less.add(k)
            elif ".downsample." in k:
                relu_less.add(k)
        for sp in specs:
            if not isinstance(sp, A.Conv):
                continue
            k = sp.key
            w = sd[k + ".weight"].float().contiguous()
            cb = sd.get(k + ".bias")
            if k in bn_after:
                scale, shift = _fold(sd, bn_after[k], cb, w.device)
            else:
                scale, shift = None, (cb.float() if cb is not None else None)
            if k == "conv1":          # 3x3x3 stem conv == 1x1 conv over the 27(+5 zero)-channel im2col slab
                w32 = torch.zeros((64, 32, 1, 1), dtype=torch.float32)
                w32[:, :27, 0, 0] = w.reshape(64, 27).cpu()
                self.layers[k] = ConvLayer(w32.to(w.device), scale, shift, stride=1, relu=True)
                continue
            leaf = k.rsplit(".", 1)[-1]
            # residual convs (block conv2 / bottleneck conv3) apply ReLU after the add -> flag set
            relu = k not in relu_less and k != "last_layer.3"
            self.layers[k] = ConvLayer(w, scale, shift, stride=sp.stride, relu=relu, out_nchw=(k == "last_layer.3"))
        if self.variant == "softmax":
            self.temp = sd["trainable_temp"].detach().float().reshape(1).clone()
        torch.cuda.synchronize(self.device)

    def plan(self, B, H, W):
        key = (B, H, W)
        if key not in self.plans:
            with torch.cuda.device(self.device), _BUILD_LOCK:
                if key not in self.plans:
                    self.plans[key] = Plan(self, B, H, W)
        return self.plans[key]

    def forward_u8(self, images, mean, std, want_features=True):
        """images: [B,H,W,3] uint8 CUDA tensor (what the data loader holds before ToTensor + Normalize); the normalisation of
        lib/dataset/transforms/build.py:82-85 is applied inside the stem's im2col kernel, so the 4x larger fp32 NCHW image
        tensor the reference uploads never exists."""
        import ctypes as C_
        if images.dim() != 4 or images.shape[3] != 3 or images.dtype != torch.uint8:
            raise ValueError("expected uint8 images [B, H, W, 3]")
        B, H, W, _ = images.shape
        p = self.plan(B, H, W)
        with torch.cuda.device(self.device):
            if p.x_u8 is None:
                p.x_u8 = torch.zeros((B, H, W, 3), dtype=torch.uint8, device=self.device)
            norm = ((C_.c_float * 3)(*[float(v) for v in mean]), (C_.c_float * 3)(*[float(v) for v in std]))
            if p.norm is not None and (list(p.norm[0]) != list(norm[0]) or list(p.norm[1]) != list(norm[1])):
                p.graph = None                      # the constants are baked into the captured launch
            p.norm = norm
            p.x_u8.copy_(images, non_blocking=True)
            return p.run(want_features, self.use_graph, u8=True)

    def forward(self, x, want_features=True):
        """x: [B,3,H,W] float32 CUDA NCHW.  Returns the plan's static output tensors (overwritten by the next
        call with the same shape - clone to keep)."""
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected input [B, 3, H, W]")
        B, _, H, W = x.shape
        p = self.plan(B, H, W)
        with torch.cuda.device(self.device):
            p.x.copy_(x, non_blocking=True)
            return p.run(want_features, self.use_graph)
