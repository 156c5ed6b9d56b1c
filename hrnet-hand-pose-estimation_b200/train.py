"""Training engine: forward with batch-statistics BatchNorm, the fused heat-map / pose2d losses, the full backward
pass (data- and weight-gradients of all 307 convs, BN / fuse / bilinear / softmax backward) and the fused Adam step,
as launch plans over libhrnb.so.  Host side only: every number is produced by the CUDA kernels.

What it replaces in the reference (file:line relative to the reference repo):
  model.train() forward            lib/models/pose_hrnet.py:511-568 / pose_hrnet_softmax.py:449-528 (BN in train mode)
  get_final_preds(use_softmax)     lib/utils/heatmap_decoding.py:87-101 (soft-argmax inside the step)
  AverageMeter.computeLosses       lib/core/function.py:1334-1344  (total = f_hm * HeatmapLoss + f_p2d * JointsMSELoss)
  total_loss.backward()            lib/core/function.py:101-106    (autograd; here hand-written kernels)
  optimizer.step() (Adam, L2 wd)   lib/utils/utils.py:71-92

A `TrainPlan` is built for one (batch, H, W): forward ops are recorded in order together with a "tape" of backward
builders; the backward launch list is generated from the tape in reverse, deciding per gradient buffer whether a
kernel writes (first contribution) or accumulates (later ones).  Each unit conv -> BN -> (+residual) -> ReLU keeps its
conv output `c` and its output `y` (bf16 PF8) for the backward pass; gradients are bf16 PF8, parameter gradients fp32.
"""
import ctypes as C
import os

import torch

from . import _lib, arch as A, tops
from .flat import FlatParams
from .ops import ConvLayer, PF8, PhasePF8, Repacker, attach_stats, grouped_conv_params


class T:
    """activation node: value + (lazily allocated) gradient buffer"""
    __slots__ = ("v", "g", "ginit", "g_sid", "bp", "bp_sid")

    def __init__(self, v, g=None):
        self.v, self.g, self.ginit, self.g_sid = v, g, False, None
        self.bp, self.bp_sid = None, None      # BatchNorm launch struct / stream of the unit that produced v (TrainPlan.unit)


def _like(v, device):
    if isinstance(v, PhasePF8):
        return PhasePF8(v.N, v.C, v.H, v.W, device=device)
    return PF8(v.N, v.C, v.H, v.W, device=device)


def _phase_view(ph, i):
    g = ph.half
    return PF8(g.N, g.C, g.H, g.W, buf=ph.buf[i])


class TrainPlan:
    def __init__(self, eng, B, H, W, feat_grad=False):
        self.eng, self.B, self.H, self.W = eng, B, H, W
        self.feat_grad = feat_grad      # the backward pass also takes a gradient w.r.t. the returned feature tensor
        self.dev = eng.device
        self.fwd, self.loss_steps, self.bwd = [], [], []      # fwd / bwd: (kind, sid, fn | other stream, name)
        self._ctx, self._sid = "", 0
        self.multi_stream = eng.multi_stream
        self.all_bufs = []      # every activation / gradient buffer of the plan: the launch parameter structs hold raw
                                # device pointers only, so the plan must own the tensors for as long as it can be replayed
        self.tape = []
        self.keep = []
        self.n_launch = {"fwd": 0, "loss": 0, "bwd": 0}
        self.act_bytes = 0
        self.x = torch.zeros((B, 3, H, W), dtype=torch.float32, device=self.dev)
        self.graphs = {}
        self.conv_out = {}      # conv key -> PF8 conv output (pre-BN) kept for the backward pass
        self.cuts = []          # (first parameter prefix whose gradients are final, index into self.bwd): all-reduce bucket cuts
        self.generation = 0     # bumped by every forward that overwrites the saved activations (models/_hrnet.py checks it)
        self._build()

    # ---- helpers ------------------------------------------------------------------------------------------
    def _buf(self, C_, H, W):
        t = PF8(self.B, C_, H, W, device=self.dev)
        self.act_bytes += t.buf.numel() * 2
        self.all_bufs.append(t)
        return t

    def _grad(self, t):
        if t.g is None:
            t.g = _like(t.v, self.dev)
            self.act_bytes += t.g.buf.numel() * 2
            self.all_bufs.append(t.g)
        return t.g

    def _f(self, fn, name=""):
        self.fwd.append(("op", self._sid, fn, name or self._ctx))
        self.n_launch["fwd"] += 1

    def _b(self, fn, name=""):
        self.bwd.append(("op", self._sid, fn, name or self._ctx))
        self.n_launch["bwd"] += 1

    def _wait(self, steps, sid, other):
        """stream `sid` waits for everything queued on stream `other` so far"""
        if self.multi_stream and sid != other and other is not None:
            steps.append(("wait", sid, other, ""))

    def on(self, sid):
        self._sid = sid if self.multi_stream else 0

    def fwait(self, sid, other):
        self._wait(self.fwd, sid, other)

    def _gr(self, t):
        """backward op on the current stream is about to READ t.g: order it after the stream that wrote it last"""
        self._wait(self.bwd, self._sid, t.g_sid)

    def _gw(self, t):
        """backward op on the current stream is about to write / accumulate into t.g -> (buffer, mode)"""
        g = self._grad(t)
        self._wait(self.bwd, self._sid, t.g_sid)
        mode = 2 if t.ginit else 1
        t.ginit, t.g_sid = True, self._sid
        return g, mode

    @property
    def fwd_fns(self):
        return [st[2] for st in self.fwd if st[0] == "op"]

    @property
    def fwd_names(self):
        return [st[3] for st in self.fwd if st[0] == "op"]

    @property
    def bwd_fns(self):
        return [st[2] for st in self.bwd if st[0] == "op"]

    @property
    def bwd_names(self):
        return [st[3] for st in self.bwd if st[0] == "op"]

    def _conv_fn(self, layer, x, out, res=None, stats=None):
        """-> launch closure; with `stats` (fp32 [cout, 2]) -> (closure, fused): fused = the conv also writes the
        BatchNorm batch statistics of `out` (eligible tile shapes only, ops.stats_eligible)"""
        layer.no_pdl = not self.eng.pdl
        want = stats is not None and self.eng.fuse_stats and res is None and layer.cout in (16, 32, 64)
        p = layer.params(x, out, res, bn=layer.cout if want else None)     # fused statistics need one N tile
        self.keep.append(p)
        lib, ref = _lib.lib(), C.byref(p)
        fn = lambda: _lib.check(lib.hrnb_conv(ref, _lib.stream_ptr()))
        if stats is None:
            return fn
        return fn, (want and attach_stats(p, stats))

    def _wgrad_fn(self, dy, x_ptr, x_ps, dw, cin, cout, taps, src_stride=0):
        p = tops.wgrad_params(dy, x_ptr, x_ps, dw, cin, cout, taps, src_stride=src_stride)
        self.keep.append(p)
        lib, ref = _lib.lib(), C.byref(p)
        return lambda: _lib.check(lib.hrnb_wgrad(ref, _lib.stream_ptr()))

    # ---- ops ----------------------------------------------------------------------------------------------
    def split(self, x):
        """PF8 -> PhasePF8 copy for the stride-2 convs reading x (one split serves all of them)"""
        ph = T(PhasePF8(self.B, x.v.C, x.v.H, x.v.W, device=self.dev))
        self.act_bytes += ph.v.buf.numel() * 2
        self.all_bufs.append(ph.v)
        lib, s, d = _lib.lib(), x.v, ph.v
        bp = getattr(x, "bp", None)
        if bp is not None and not bp.out2 and getattr(x, "bp_sid", None) == self._sid and os.environ.get("HRNB_SPLIT_KERNEL", "0") != "1":
            # x comes out of a unit's BatchNorm kernel on this stream: that kernel writes the phase-split copy as well
            # (hrnb_bn_params.out2) - no separate pass over the tensor
            bp.out2, bp.out2_ps, bp.out2_phase_stride = d.ptr, d.ps, d.phase_stride
        else:
            self._f(lambda: _lib.check(lib.hrnb_phase_split(s.ptr, s.ps, s.N, s.C, s.H, s.W, d.ptr, d.ps, d.phase_stride,
                                                            _lib.stream_ptr())), "phase_split")

        sid = self._sid

        def back():
            assert ph.ginit
            self._ctx, self._sid = "phase_merge", sid
            self._gr(ph)
            g = ph.g
            dst, mode = self._gw(x)
            self._b(lambda: tops.phase_merge(g, dst, mode), "phase_merge")
        self.tape.append(back)
        return ph

    def unit(self, key, x, relu, res=None, need_dx=True):
        """conv `key` -> BatchNorm (batch statistics) -> (+ res) -> (ReLU); x: T of PF8 (stride 1) / PhasePF8 (stride 2)"""
        e = self.eng
        L = e.units[key]
        sp = L["spec"]
        stride, k = sp.stride, sp.k
        Ho, Wo = x.v.H // stride, x.v.W // stride
        self._ctx = key
        c = self._buf(sp.cout, Ho, Wo)
        y = T(self._buf(sp.cout, Ho, Wo))
        sums, dsums = L["sums"], L["dsums"]
        conv_fn, fused = self._conv_fn(L["fwd"], x.v, c, stats=sums)
        self._f(conv_fn, "conv:" + key)
        self.conv_out[key] = c
        sid = self._sid
        bp = tops.bn_params(c, sums, L["gamma"], L["beta"], y.v, res=res.v if res is not None else None, relu=relu,
                            running_mean=L["rm"], running_var=L["rv"])
        self.keep.append(bp)
        lib, bref = _lib.lib(), C.byref(bp)
        if not fused:      # else: the conv's epilogue already reduced the batch statistics (conv_tc.cu, STATS variant)
            self._f(lambda: tops.bn_stats(c, sums, sid), "bn_stats:" + key)
        self._f(lambda: _lib.check(lib.hrnb_bn_apply(bref, _lib.stream_ptr())), "bn_apply:" + key)
        y.bp, y.bp_sid = bp, sid          # split() may ask this launch for a phase-split copy of y

        def back():
            assert y.ginit, key
            self._ctx, self._sid = key, sid
            self._gr(y)
            dy = y.g
            dres, dmode = None, 0
            if res is not None:
                dres, dmode = self._gw(res)
            bb = tops.bn_bwd_params(dy, y.v, c, sums, L["gamma"], dsums, dy, L["dgamma"], L["dbeta"], relu=relu,
                                    dres=dres, dres_mode=dmode, sid=sid)
            self.keep.append(bb)
            r = C.byref(bb)
            self._b(lambda: _lib.check(lib.hrnb_bn_bwd_reduce(r, _lib.stream_ptr())), "bn_bwd_reduce:" + key)
            self._b(lambda: _lib.check(lib.hrnb_bn_bwd_apply(r, _lib.stream_ptr())), "bn_bwd_apply:" + key)
            self._conv_backward(L, x, dy, need_dx)
        self.tape.append(back)
        return y

    def units(self, keys, xs, relu, ress=None):
        """n <= 4 independent units (the branches of a module at the same depth): convs launched one by one, their
        BatchNorm statistics / normalisation and the BatchNorm backward horizontally batched into two launches each"""
        e, lib, n = self.eng, _lib.lib(), len(keys)
        ress = ress or [None] * n
        Ls, cs, ys, bps = [], [], [], []
        have_stats = 0
        for j, (key, x, res) in enumerate(zip(keys, xs, ress)):
            L = e.units[key]
            sp = L["spec"]
            assert sp.stride == 1
            self._ctx = key
            c = self._buf(sp.cout, x.v.H, x.v.W)
            y = T(self._buf(sp.cout, x.v.H, x.v.W))
            conv_fn, fused = self._conv_fn(L["fwd"], x.v, c, stats=L["sums"])
            have_stats |= int(bool(fused)) << j
            self._f(conv_fn, "conv:" + key)
            self.conv_out[key] = c
            bps.append(tops.bn_params(c, L["sums"], L["gamma"], L["beta"], y.v, res=res.v if res is not None else None,
                                      relu=relu, running_mean=L["rm"], running_var=L["rv"]))
            Ls.append(L); cs.append(c); ys.append(y)
        arr = (_lib.BnParams * n)(*bps)
        sid = self._sid
        ws = tops.reduce_ws(self.dev, sid)
        self.keep += [arr, bps]
        self._f(lambda: _lib.check(lib.hrnb_bn_forward_batch(arr, n, ws.data_ptr(), have_stats, _lib.stream_ptr())),
                "bn_fwd_batch:" + keys[0])
        if have_stats != (1 << n) - 1:
            self.n_launch["fwd"] += 1       # statistics launch for the tensors the convs did not cover + normalisation launch

        def back():
            self._ctx, self._sid = keys[0], sid
            bbs = []
            for L, c, y, res in zip(Ls, cs, ys, ress):
                assert y.ginit
                self._gr(y)
                dres, dmode = (None, 0) if res is None else self._gw(res)
                bbs.append(tops.bn_bwd_params(y.g, y.v, c, L["sums"], L["gamma"], L["dsums"], y.g, L["dgamma"], L["dbeta"],
                                              relu=relu, dres=dres, dres_mode=dmode, sid=sid))
            barr = (_lib.BnBwdParams * n)(*bbs)
            self.keep += [barr, bbs]
            self._b(lambda: _lib.check(lib.hrnb_bn_backward_batch(barr, n, _lib.stream_ptr())), "bn_bwd_batch:" + keys[0])
            self.n_launch["bwd"] += 1
            for L, x, y in zip(Ls, xs, ys):
                self._ctx = L["spec"].key
                self._conv_backward(L, x, y.g, True)
        self.tape.append(back)
        return ys

    def _bw(self, fn, name):
        """record a weight-gradient launch.  Nothing downstream in the backward pass reads dW, so with the multi-stream plan
        it goes to the companion stream (4 + s) of the branch stream s, ordered after the kernel that produced dc: the
        chain bn_bwd -> dgrad -> bn_bwd ... of the branch no longer waits for its weight gradients."""
        if not (self.multi_stream and self.eng.wgrad_streams):
            return self._b(fn, name)
        sid = self._sid
        wsid = 4 if self.eng.wgrad_streams == 2 else 4 + sid      # 2: one shared weight-gradient stream
        if not self._wg_forked:        # one fork per producer of dc, however many wgrad launches follow
            self._wait(self.bwd, wsid, sid)
            self._wg_forked = True
        self._sid = wsid
        self._b(fn, name)
        self._sid = sid

    def _conv_backward(self, L, x, dc, need_dx):
        """weight gradient of conv L from (dc, x) and, if asked, the data gradient into x.g"""
        sp = L["spec"]
        dw = L["dw"]
        cin_g = dw.shape[1]
        self._wg_forked = False
        if sp.stride == 1:
            self._bw(self._wgrad_fn(dc, x.v.ptr, x.v.ps, dw, cin_g, sp.cout, tops.fwd_taps_s1(sp.k, dc.Wp)), "wgrad:" + sp.key)
            if need_dx:
                gx, mode = self._gw(x)
                self._b(self._conv_fn(L["dgrad"], dc, gx, res=gx if mode == 2 else None), "dgrad:" + sp.key)
        else:
            if os.environ.get("HRNB_S2_SPLIT", "0") == "1":       # A/B: one weight-gradient launch per input phase (round-2 first half)
                for ph, taps in tops.fwd_taps_s2(dc.Wp).items():
                    self._bw(self._wgrad_fn(dc, x.v.ptr + ph * x.v.phase_stride * 2, x.v.ps, dw, cin_g, sp.cout, taps), "wgrad:" + sp.key)
            else:
                self._bw(self._wgrad_fn(dc, x.v.ptr, x.v.ps, dw, cin_g, sp.cout, tops.fwd_taps_s2_merged(dc.Wp),
                                        src_stride=x.v.phase_stride), "wgrad:" + sp.key)
            if need_dx:
                gx, mode = self._gw(x)
                outs = [_phase_view(gx, ph) for ph in range(4)]
                self.keep.extend(outs)
                if os.environ.get("HRNB_S2_SPLIT", "0") == "1":       # A/B: one data-gradient launch per input phase
                    for ph in range(4):
                        self._b(self._conv_fn(L["dgrad"][ph], dc, outs[ph], res=outs[ph] if mode == 2 else None), "dgrad:" + sp.key)
                else:
                    # the four per-phase convs (1, 2, 2 and 4 taps) as ONE grouped launch (include/hrnb.h: ngroup)
                    layers = [L["dgrad"][ph] for ph in range(4)]
                    for l in layers:
                        l.no_pdl = not self.eng.pdl
                    gp = grouped_conv_params(layers, dc, outs, outs if mode == 2 else None)
                    self.keep.append(gp)
                    lib, ref = _lib.lib(), C.byref(gp)
                    self._b(lambda lib=lib, ref=ref: _lib.check(lib.hrnb_conv(ref, _lib.stream_ptr())), "dgrad:" + sp.key)

    def fuse(self, srcs, shifts, out_v=None, out_g=None):
        ch, (H, W) = srcs[0].v.C, (srcs[0].v.H << shifts[0], srcs[0].v.W << shifts[0])
        out = T(out_v if out_v is not None else self._buf(ch, H, W), out_g)
        p = _lib.FuseParams()
        for i, (s, sh) in enumerate(zip(srcs, shifts)):
            p.src[i], p.src_ps[i], p.shift[i] = s.v.ptr, s.v.ps, sh
        p.nsrc = len(srcs)
        p.out, p.out_ps = out.v.ptr, out.v.ps
        p.N, p.H, p.W, p.C, p.relu = self.B, H, W, ch, 1
        self.keep.append(p)
        lib, ref = _lib.lib(), C.byref(p)
        self._f(lambda: _lib.check(lib.hrnb_fuse_sum(ref, _lib.stream_ptr())), "fuse")

        sid = self._sid

        def back():
            assert out.ginit
            self._ctx, self._sid = "fuse", sid
            self._gr(out)
            # one launch per source by default: the batched form (hrnb_fuse_sum_bwd_batch, HRNB_FUSE_BWD_BATCH=1) saves 62 launches
            # but has to wait for the gradient streams of ALL sources at once - measured neutral (21.33 vs 21.34 ms/step)
            if os.environ.get("HRNB_FUSE_BWD_BATCH", "0") != "1":
                for s, sh in zip(srcs, shifts):
                    g, mode = self._gw(s)
                    self._b(lambda g=g, sh=sh, mode=mode: tops.fuse_sum_bwd(out.g, out.v, g, sh, True, mode), "fuse_bwd")
            else:
                gm = [self._gw(s) for s in srcs]
                args = tops.fuse_sum_bwd_batch_args(out.g, out.v, [g for g, _ in gm], list(shifts), [m for _, m in gm], True)
                self.keep.append(args)
                self._b(lambda args=args: _lib.check(lib.hrnb_fuse_sum_bwd_batch(*args, _lib.stream_ptr())), "fuse_bwd")
        self.tape.append(back)
        return out

    # ---- the network --------------------------------------------------------------------------------------
    def _build(self):
        e, B = self.eng, self.B
        arch, lib = e.arch, _lib.lib()
        ch = arch.channels
        if self.H % 32 or self.W % 32:
            raise ValueError("input H and W must be multiples of 32 (four resolutions, each halving)")
        H2, W2, H4, W4 = self.H // 2, self.W // 2, self.H // 4, self.W // 4
        J = arch.num_joints

        # per-step zeroing of the statistics workspace and the parameter gradients
        self._f(lambda: e.stats.zero_(), "zero")
        self._f(lambda: e.flat.grads.zero_(), "zero")
        self._f(lambda: torch._foreach_add_(e.nbt, 1), "zero")

        cols = T(self._buf(32, H2, W2))
        x = self.x
        self._f(lambda: _lib.check(lib.hrnb_stem_im2col(x.data_ptr(), cols.v.ptr, cols.v.ps, B, self.H, self.W,
                                                        _lib.stream_ptr())), "stem_im2col")
        t = self.unit("conv1", cols, True, need_dx=False)
        cur = self.unit("conv2", self.split(t), True)

        for b in range(4):
            pre = "layer1.%d" % b
            c1 = self.unit(pre + ".conv1", cur, True)
            c2 = self.unit(pre + ".conv2", c1, True)
            res = self.unit(pre + ".downsample.0", cur, False) if b == 0 else cur
            cur = self.unit(pre + ".conv3", c2, True, res=res)

        # branch i of the multi-resolution stages lives on stream i (they only meet in the fuse layers)
        xs = [self.unit("transition1.0.0", cur, True)]
        ph = self.split(cur)
        self.fwait(1, 0)
        self.on(1)
        xs.append(self.unit("transition1.1.0.0", ph, True))
        stage3_b0 = None
        cat = T(None)
        for s, nmod in zip((2, 3, 4), arch.modules):
            nb = s
            if s > 2:
                # everything recorded on the tape AFTER this point (transition{s-1}, stage s, ..., head) has its parameter
                # gradients final once the backward pass comes back here: a cut for the bucketed gradient all-reduce
                self.tape.append(lambda tag="transition%d." % (s - 1): self.cuts.append((tag, len(self.bwd))))
                self.on(nb - 2)
                ph = self.split(xs[-1])
                self.fwait(nb - 1, nb - 2)
                self.on(nb - 1)
                xs.append(self.unit("transition%d.%d.0.0" % (s - 1, nb - 1), ph, True))
            for m in range(nmod):
                pre = "stage%d.%d" % (s, m)
                last_module = (s == 4 and m == nmod - 1)
                splits = {}
                if e.bn_batch and not self.multi_stream:
                    # single-stream plan: walk the branches in lock-step so that their BatchNorm kernels batch horizontally
                    for b in range(arch.blocks):
                        bps_ = ["%s.branches.%d.%d" % (pre, i, b) for i in range(nb)]
                        ys_ = self.units([p_ + ".conv1" for p_ in bps_], xs, True)
                        xs = self.units([p_ + ".conv2" for p_ in bps_], ys_, True, ress=xs)
                    for i in range(nb - 1):
                        splits[i] = self.split(xs[i])
                else:
                    for i in range(nb):
                        self.on(i)
                        for b in range(arch.blocks):
                            bp = "%s.branches.%d.%d" % (pre, i, b)
                            y = self.unit(bp + ".conv1", xs[i], True)
                            xs[i] = self.unit(bp + ".conv2", y, True, res=xs[i])
                        if i < nb - 1:
                            splits[i] = self.split(xs[i])      # phase copy for the stride-2 chains that start at branch i
                for i in range(nb):                        # every fuse output needs every branch
                    for j in range(nb):
                        self.fwait(i, j)
                outs = []
                for i in range(nb):
                    self.on(i)
                    srcs, shifts = [], []
                    # the 1x1 convs from the lower-resolution branches are independent units on this stream; batching their
                    # BatchNorm kernels horizontally (HRNB_FUSE_BN_BATCH=1: -39 launches) measured SLOWER in-trip (20.91 vs 20.75
                    # ms/step): conv -> BN chains of separate units overlap through PDL, the batched form waits for all convs
                    up_keys = ["%s.fuse_layers.%d.%d.0" % (pre, i, j) for j in range(i + 1, nb)]
                    ups_ = None
                    if len(up_keys) >= 2 and os.environ.get("HRNB_FUSE_BN_BATCH", "0") == "1":
                        ups_ = self.units(up_keys, [xs[j] for j in range(i + 1, nb)], False)
                    for j in range(nb):
                        if j == i:
                            srcs.append(xs[j]); shifts.append(0)
                        elif j > i:
                            srcs.append(ups_[j - i - 1] if ups_ is not None else
                                        self.unit("%s.fuse_layers.%d.%d.0" % (pre, i, j), xs[j], False))
                            shifts.append(j - i)
                        else:
                            t = splits[j]
                            for k in range(i - j):
                                last = k == i - j - 1
                                t = self.unit("%s.fuse_layers.%d.%d.%d.0" % (pre, i, j, k), t, not last)
                                if not last:
                                    t = self.split(t)
                            srcs.append(t); shifts.append(0)
                    if last_module and i == 0:
                        cat.v = self._buf(arch.head_channels, H4, W4)
                        cat.g = self._grad(cat)
                        outs.append(self.fuse(srcs, shifts, cat.v.view_planes(0, ch[0] // 8), cat.g.view_planes(0, ch[0] // 8)))
                        cat_b0 = outs[-1]
                    else:
                        outs.append(self.fuse(srcs, shifts))
                xs = outs
            if s == 3:
                stage3_b0 = xs[0]

        # head: bilinear up-sampling of branches 1..3 into the concat buffer (each on its branch's stream)
        align = 1 if e.variant == "softmax" else 0
        plane0 = ch[0] // 8
        ups = []
        for i in range(1, 4):
            self.on(i)
            dst = cat.v.view_planes(plane0, ch[i] // 8)
            gdst = cat.g.view_planes(plane0, ch[i] // 8)
            plane0 += ch[i] // 8
            src = xs[i]
            self.keep += [dst, gdst]
            self._f(lambda src=src, dst=dst: _lib.check(lib.hrnb_bilinear_up(
                src.v.ptr, src.v.ps, src.v.N, src.v.C, src.v.H, src.v.W, dst.ptr, dst.ps, dst.H, dst.W, align, _lib.stream_ptr())),
                "bilinear")
            ups.append((i, src, gdst))
        for i in range(1, 4):
            self.fwait(0, i)
        self.on(0)

        def back_head():
            assert cat.ginit
            cat_b0.ginit, cat_b0.g_sid = True, cat.g_sid
            for i, src, gdst in ups:
                self.on(i)
                self._gr(cat)
                g, mode = self._gw(src)
                self._b(lambda g=g, gdst=gdst, mode=mode: tops.bilinear_up_bwd(gdst, g, align, mode), "bilinear_bwd")
        self.tape.append(back_head)

        hid = self.unit("last_layer.0", cat, True)

        # final conv (+bias, no BN) -> fp32 NCHW logits
        F3 = e.final
        logits = torch.empty((B, J, H4, W4), dtype=torch.float32, device=self.dev)
        self._f(self._conv_fn(F3["fwd"], hid.v, logits), "conv:last_layer.3")
        d_logits = torch.zeros((B, J, H4, W4), dtype=torch.float32, device=self.dev)
        dlog = self._buf(32 if J <= 32 else (J + 15) // 16 * 16, H4, W4)

        def back_final():
            self.on(0)
            self._ctx = "last_layer.3"
            self._b(lambda: _lib.check(lib.hrnb_nchw_f32_to_pf8(d_logits.data_ptr(), B, J, H4, W4, dlog.ptr, dlog.ps,
                                                                _lib.stream_ptr())), "nchw_to_pf8")
            if F3["dbias"] is not None:
                self._b(lambda: tops.channel_sum(dlog, F3["dbias"], J, 0), "channel_sum")
            k = F3["spec"].k
            self._b(self._wgrad_fn(dlog, hid.v.ptr, hid.v.ps, F3["dw"], arch.head_channels, J, tops.fwd_taps_s1(k, dlog.Wp)), "wgrad:last_layer.3")
            gx, _ = self._gw(hid)
            self._b(self._conv_fn(F3["dgrad"], dlog, gx), "dgrad:last_layer.3")
        self.tape.append(back_final)
        self.out = {"logits": logits}
        self.d_logits = d_logits
        self.cat, self.stage3_b0 = cat, stage3_b0

        # decode + losses (fused path): softmax -> heat-map + soft-argmax -> HeatmapLoss + JointsMSELoss -> d_logits
        self.gt_heat = torch.zeros((B, J, H4, W4), dtype=torch.float32, device=self.dev)
        self.gt_xy = torch.zeros((B, J, 2), dtype=torch.float32, device=self.dev)
        self.vis = torch.ones((B, J), dtype=torch.float32, device=self.dev)
        self.losses = torch.zeros(3, dtype=torch.float32, device=self.dev)     # total, heat-map, pose2d
        ws = torch.empty(1024, dtype=torch.float32, device=self.dev)
        f_hm, f_p2d = e.loss_factors
        hm_scale = torch.tensor([f_hm], dtype=torch.float32, device=self.dev)
        BJ = B * J
        if e.variant == "softmax":
            heat = torch.empty_like(logits)
            coords = torch.empty((B, J, 2), dtype=torch.float32, device=self.dev)
            self.d_heat = torch.zeros_like(logits)
            self.d_coords = torch.zeros_like(coords)
            temp = e.temp_param
            self._f(lambda: _lib.check(lib.hrnb_softmax_softargmax(logits.data_ptr(), temp.data_ptr(), BJ, H4, W4,
                                                                   heat.data_ptr(), coords.data_ptr(), _lib.stream_ptr())),
                    "softmax_softargmax")
            self.out["heatmap"], self.out["coords"] = heat, coords

            def loss_fused():
                _lib.check(lib.hrnb_loss_heatmap(heat.data_ptr(), self.gt_heat.data_ptr(), BJ, H4 * W4, 0,
                                                 self.losses[1:2].data_ptr(), self.d_heat.data_ptr(), hm_scale.data_ptr(),
                                                 ws.data_ptr(), _lib.stream_ptr()))
                _lib.check(lib.hrnb_loss_pose2d(coords.data_ptr(), self.gt_xy.data_ptr(), self.vis.data_ptr(), B, J,
                                                self.losses[2:3].data_ptr(), self.d_coords.data_ptr(), _lib.stream_ptr()))
                self.d_coords.mul_(f_p2d)
                torch.add(self.losses[1] * f_hm, self.losses[2], alpha=f_p2d, out=self.losses[0])
            self.loss_steps.append(loss_fused)
            self.n_launch["loss"] += 5

            dtemp = e.dtemp

            def softmax_back():
                _lib.check(lib.hrnb_softmax_softargmax_bwd(logits.data_ptr(), temp.data_ptr(), heat.data_ptr(),
                                                           self.d_heat.data_ptr(), self.d_coords.data_ptr(), BJ, H4, W4,
                                                           d_logits.data_ptr(), dtemp.data_ptr() if dtemp is not None else None,
                                                           _lib.stream_ptr()))
            self.softmax_back = softmax_back
        else:
            def loss_fused():
                _lib.check(lib.hrnb_loss_heatmap(logits.data_ptr(), self.gt_heat.data_ptr(), BJ, H4 * W4, 0,
                                                 self.losses[1:2].data_ptr(), d_logits.data_ptr(), hm_scale.data_ptr(),
                                                 ws.data_ptr(), _lib.stream_ptr()))
                torch.mul(self.losses[1], f_hm, out=self.losses[0])
            self.loss_steps.append(loss_fused)
            self.n_launch["loss"] += 3
            self.softmax_back = None

        # feature output (NCHW fp32) as the reference returns it
        feat_src = cat.v if e.variant == "softmax" else stage3_b0.v
        self.feat = torch.empty((B, feat_src.C, feat_src.H, feat_src.W), dtype=torch.float32, device=self.dev)
        self.feat_fn = lambda: _lib.check(lib.hrnb_pf8_to_nchw_f32(feat_src.ptr, feat_src.ps, feat_src.N, feat_src.C,
                                                                   feat_src.H, feat_src.W, self.feat.data_ptr(),
                                                                   _lib.stream_ptr()))

        # gradient w.r.t. the feature output (inter_feat): the second return value of the reference's forward is an
        # autograd-connected tensor (pose_hrnet_softmax.py:528 the 480-channel concat, pose_hrnet.py:568 stage-3 branch 0) and
        # wrappers train through it (pose_hrnet_volumetric.py:620-634).  It is injected FIRST in the backward pass as the
        # initial content of that activation's gradient buffer; every later contribution then accumulates.
        self.d_feat = None
        if self.feat_grad:
            ft = cat if e.variant == "softmax" else stage3_b0
            self.d_feat = torch.zeros((B, feat_src.C, feat_src.H, feat_src.W), dtype=torch.float32, device=self.dev)
            d_feat = self.d_feat

            def back_feat():
                self.on(0)
                self._ctx = "inter_feat"
                g, mode = self._gw(ft)
                assert mode == 1
                self._b(lambda: _lib.check(lib.hrnb_nchw_f32_to_pf8(d_feat.data_ptr(), B, g.C, g.H, g.W, g.ptr, g.ps,
                                                                    _lib.stream_ptr())), "nchw_to_pf8:inter_feat")
            self.tape.append(back_feat)

        # backward launch list from the tape
        self.on(0)
        if self.softmax_back is not None:
            self._b(self.softmax_back, "softmax_bwd")
        for back in reversed(self.tape):
            back()
        self.tape = None

    # ---- execution ------------------------------------------------------------------------------------------
    def _run(self, steps):
        lib = _lib.lib()
        lib.hrnb_debug_set(4, 1 if self.eng.pdl else 0)     # PDL attribute for the elementwise / wgrad launches of THIS plan only
        try:
            self._run_steps(steps)
        finally:
            lib.hrnb_debug_set(4, 0)

    def _run_steps(self, steps):
        main = torch.cuda.current_stream()
        side = self.eng.side_streams if self.multi_stream else []
        streams = [main] + side
        for sd in side:
            sd.wait_stream(main)
        for kind, sid, a, _ in steps:
            if kind == "op":
                if sid == 0:
                    a()
                else:
                    with torch.cuda.stream(streams[sid]):
                        a()
            else:
                streams[sid].wait_stream(streams[a])
        for sd in side:
            main.wait_stream(sd)

    def run_forward(self, want_features=False):
        self._run(self.fwd)
        if want_features:
            self.feat_fn()

    def run_loss(self):
        for fn in self.loss_steps:
            fn()

    def run_backward(self, lo=0, hi=None):
        self._run(self.bwd[lo:hi])

    def ar_segments(self):
        """backward list split at the all-reduce cuts -> [(bwd lo, bwd hi, grad lo, grad hi)]: once bwd[lo:hi] has run, the
        gradient elements [grad lo, grad hi) of the flat buffer are final (the buffer is in named_parameters() order: stem ...
        stage4, head; the backward pass walks it from the end)"""
        flat, names = self.eng.flat, self.eng.param_names
        out, prev_b, prev_g = [], 0, flat.n_grads
        for tag, bidx in self.cuts:                 # in backward order: "transition3." first, then "transition2."
            first = min(i for i, n in enumerate(names) if n.startswith(tag))
            g0 = flat.g_offs[first]
            out.append((prev_b, bidx, g0, prev_g))
            prev_b, prev_g = bidx, g0
        out.append((prev_b, len(self.bwd), 0, prev_g))
        return out

    def _graphed(self, name, body):
        if not self.eng.use_graph:
            body()
            return
        if name not in self.graphs:
            body()                       # warm-up outside capture (function attributes, lazy module loads)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread may poll events of an in-flight gradient all-reduce during the capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                body()
            self.graphs[name] = g
            return                       # the warm-up run already produced this step's results ... replay for state parity
        self.graphs[name].replay()


class TrainEngine:
    """Owns the flat parameter buffers, the per-layer packed weights (forward and data-gradient orientation) and the
    per-shape TrainPlans of one network on one device."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, loss_factors=(1.0, 0.1),
                 use_graph=True, multi_stream=None, bn_batch=None):
        self.model = model
        self.arch, self.variant = model.arch, model.variant
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("the B200 HRNet trains on CUDA only (no CPU fallback): call .cuda() first")
        self.loss_factors = tuple(float(f) for f in loss_factors)
        self.use_graph = use_graph and os.environ.get("HRNB_NO_GRAPH", "0") != "1"
        # Default: branches of a HighResolutionModule on parallel streams, plain stream-ordered launches.
        # HRNB_TRAIN_STREAMS=0: single-stream plan with the BatchNorm kernels of a module's branches batched horizontally.
        # Programmatic dependent launch (PDL) on every kernel of the step is the default since round 2: 22.98 vs 24.83 ms/step at
        # batch 64.  Round 1 kept it opt-in because about one PDL run in eight ended in a device-side mbarrier time-out; the
        # cause (conv_tc.cu producer: the prefetched weight stage must not wait on empty_b) was fixed then, and the soak a default
        # needs was done in round 2: 8 of 8 bench runs + the whole GPU suite under PDL (profiles/r2_pdl_soak.txt).
        # HRNB_TRAIN_PDL=0 turns it off.
        self.multi_stream = os.environ.get("HRNB_TRAIN_STREAMS", "1") != "0" if multi_stream is None else bool(multi_stream)
        self.pdl = os.environ.get("HRNB_TRAIN_PDL", "1") != "0"
        # BatchNorm batch statistics reduced in the epilogue of the producing conv where the tile shape allows it (cout = BN
        # in {16, 32, 64}: the high-resolution layers); HRNB_FUSE_STATS=0: always the separate bn_stats pass
        self.fuse_stats = os.environ.get("HRNB_FUSE_STATS", "1") != "0"
        # Weight-gradient launches of the multi-stream plan run on a companion stream (nothing downstream in the backward
        # pass reads dW): HRNB_WGRAD_STREAMS=2 (default) one shared stream, 1 one per branch, 0 in line on the branch stream.
        # Measured at batch 64: 25.0 (shared) / 25.9 (in line) ms/step; one stream per branch was not faster than in line.
        self.wgrad_streams = int(os.environ.get("HRNB_WGRAD_STREAMS", "2"))
        # single-stream plan: BatchNorm kernels of the branches of a module batched horizontally (HRNB_BN_BATCH=0: off)
        self.bn_batch = os.environ.get("HRNB_BN_BATCH", "1") != "0" if bn_batch is None else bool(bn_batch)
        self.plans = {}
        with torch.cuda.device(self.device):
            _lib.hang_init()
            self.side_streams = [torch.cuda.Stream(device=self.device) for _ in range(7)]   # 1-3: branches, 4-7: their wgrads
            self._setup(lr, betas, eps, weight_decay)

    def _setup(self, lr, betas, eps, weight_decay):
        model, dev = self.model, self.device
        named = model.engine_parameters() if hasattr(model, "engine_parameters") else list(model.named_parameters())
        index = {n: i for i, (n, _) in enumerate(named)}
        self.param_names = [n for n, _ in named]
        specs = A.layer_specs(self.arch)
        convs = [sp for sp in specs if isinstance(sp, A.Conv)]
        conv_meta = {}
        for sp in convs:
            i = index[sp.key + ".weight"]
            if sp.key == "conv1":
                conv_meta[i] = (64, 27, 32, 1)
            else:
                conv_meta[i] = (sp.cout, sp.cin, sp.cin, sp.k * sp.k)
        self.flat = FlatParams([p for _, p in named], conv_meta, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        params = dict(named)           # .data now views of the flat buffer
        buffers = dict(model.named_buffers())
        self.nbt = [b for n, b in buffers.items() if n.endswith("num_batches_tracked")]
        bn_after = {}
        for a, b in zip(specs[:-1], specs[1:]):
            if isinstance(a, A.Conv) and isinstance(b, A.BN):
                bn_after[a.key] = b
        total_c = sum(b.ch for b in bn_after.values())
        self.stats = torch.zeros(total_c * 4, dtype=torch.float32, device=dev)    # per BN: sums [C][2] + dsums [C][2]
        self.repacker = Repacker(dev)
        self.units = {}
        off = 0
        gv = self.flat.grad_view
        for sp in convs:
            w = params[sp.key + ".weight"].data
            bias = params[sp.key + ".bias"].data if sp.bias else None
            if sp.key == "conv1":
                w4 = w.view(64, 27, 1, 1)
                fwd = ConvLayer(w4, None, bias, stride=1, cin_pad=32, repacker=self.repacker)
                spec = A.Conv(sp.key, 32, 64, 1, 1, False)
                dgrad = None
            else:
                spec = sp
                out_nchw = sp.key not in bn_after
                fwd = ConvLayer(w, None, bias, stride=sp.stride, out_nchw=out_nchw, repacker=self.repacker)
                if sp.stride == 1:
                    tap_ids, _ = tops.dgrad_taps_s1(sp.k, 0)
                    cpad = 32 if (out_nchw and sp.cout <= 32) else (sp.cout + 15) // 16 * 16
                    dgrad = ConvLayer(w, transpose=True, tap_ids=tap_ids, cin_pad=cpad, repacker=self.repacker)
                else:
                    dgrad = None       # geometry depends on the resolution: built per plan (see dgrad_s2)
            d = {"spec": spec, "fwd": fwd, "dgrad": dgrad, "dw": gv(index[sp.key + ".weight"]), "w": w}
            if sp.key in bn_after:
                bn = bn_after[sp.key]
                d["gamma"], d["beta"] = params[bn.key + ".weight"].data, params[bn.key + ".bias"].data
                d["dgamma"], d["dbeta"] = gv(index[bn.key + ".weight"]), gv(index[bn.key + ".bias"])
                d["rm"], d["rv"] = buffers[bn.key + ".running_mean"], buffers[bn.key + ".running_var"]
                d["sums"] = self.stats[off:off + 2 * bn.ch]
                d["dsums"] = self.stats[off + 2 * bn.ch:off + 4 * bn.ch]
                off += 4 * bn.ch
                self.units[sp.key] = d
            else:
                d["dbias"] = gv(index[sp.key + ".bias"]) if sp.bias else None
                self.final = d
        if self.variant == "softmax":
            self.temp_param = params["trainable_temp"].data.view(1)
            tp = dict(named)["trainable_temp"]
            self.dtemp = gv(index["trainable_temp"]).view(1) if tp.requires_grad else None
        self._s2_cache = {}

    def dgrad_s2(self, key, Wp_half):
        """the four per-phase data-gradient convs of stride-2 conv `key` on a half grid of padded width Wp_half"""
        ck = (key, Wp_half)
        if ck not in self._s2_cache:
            w = self.units[key]["w"]
            layers = {}
            for ph, (tap_ids, taps) in tops.dgrad_taps_s2(Wp_half).items():
                layers[ph] = ConvLayer(w, transpose=True, tap_ids=tap_ids, custom_taps=taps, repacker=self.repacker)
            self._s2_cache[ck] = layers
        return self._s2_cache[ck]

    def plan(self, B, H, W, feat_grad=False):
        key = (B, H, W) if not feat_grad else (B, H, W, True)
        if key not in self.plans:
            with torch.cuda.device(self.device):
                self._prepare_s2(H, W)     # stride-2 units: per-resolution data-gradient layers
                p = TrainPlan(self, B, H, W, feat_grad=feat_grad)
                self.repacker.run()
                torch.cuda.synchronize(self.device)
                self.plans[key] = p
        return self.plans[key]

    def _prepare_s2(self, H, W):
        """resolve the `dgrad` entry of every stride-2 unit for this input size"""
        H4, W4 = H // 4, W // 4
        for key, d in self.units.items():
            sp = d["spec"]
            if sp.stride != 2:
                continue
            if key == "conv2":
                wo = W4
            elif key.startswith("transition"):
                br = int(key.split(".")[1])
                wo = W4 >> br
            else:                                   # stageS.M.fuse_layers.I.J.K.0 : output width of chain step K
                parts = key.split(".")
                j, k = int(parts[4]), int(parts[5])
                wo = W4 >> (j + k + 1)
            d["dgrad"] = self.dgrad_s2(key, wo + 1)

    # ---- public steps ---------------------------------------------------------------------------------------
    def _touch_model(self):
        """every train-mode forward rewrites the BatchNorm running statistics through raw pointers (and graph replays do not
        bump tensor versions): tell the module so that a following model.eval() forward re-folds its weights"""
        self.model._train_epoch = getattr(self.model, "_train_epoch", 0) + 1

    def repack(self):
        self.repacker.run()

    def forward(self, x, want_features=False, feat_grad=False):
        B, _, H, W = x.shape
        p = self.plan(B, H, W, feat_grad=feat_grad)
        p.generation += 1
        self._touch_model()
        p.x.copy_(x, non_blocking=True)
        p._graphed("fwd%d" % int(want_features), lambda: p.run_forward(want_features))
        return p

    def backward(self, p):
        """d_logits (raw) or d_heat / d_coords (softmax) must be in the plan's buffers"""
        p._graphed("bwd", p.run_backward)

    def train_step(self, x, gt_heat, gt_xy=None, vis=None, optimizer_step=True, allreduce=None):
        """One fused training step: forward, losses, backward, (gradient all-reduce), Adam, weight re-pack.
        Returns the plan (losses in plan.losses = [total, heat-map, pose2d])."""
        B, _, H, W = x.shape
        p = self.plan(B, H, W)
        p.generation += 1
        self._touch_model()
        p.x.copy_(x, non_blocking=True)
        p.gt_heat.copy_(gt_heat, non_blocking=True)
        if gt_xy is not None:
            p.gt_xy.copy_(gt_xy, non_blocking=True)
        if vis is not None:
            p.vis.copy_(vis, non_blocking=True)

        if allreduce is not None and getattr(allreduce, "overlap", False):
            # data-parallel step: the backward pass is replayed in segments (head + stage 4 | stage 3 | the rest); the
            # gradient bucket a segment completes is all-reduced on the communication stream while the next segment runs
            # (what DDP's bucketed hooks do for the reference, tools/train.py:239-244)
            segs = p.ar_segments()
            for si, (blo, bhi, glo, ghi) in enumerate(segs):
                def body(si=si, blo=blo, bhi=bhi):
                    if si == 0:
                        p.run_forward(False)
                        p.run_loss()
                    p.run_backward(blo, bhi)
                p._graphed("seg%d" % si, body)
                allreduce.launch(self.flat.grads, glo, ghi)
            allreduce.finish()
        else:
            def body():
                p.run_forward(False)
                p.run_loss()
                p.run_backward()
            p._graphed("step", body)
            if allreduce is not None:
                allreduce(self.flat.grads)
        if optimizer_step:
            def opt():
                self.flat.adam_step()
                self.repacker.run()
            p._graphed("opt", opt)
            self.model._engine = None          # folded inference weights are stale now
        return p

    def launches_per_step(self, p, optimizer_step=True):
        """kernel launches of libhrnb.so in one train_step (CUDA-graph replays do not pass through the library's counter)"""
        return p.n_launch["fwd"] - 3 + p.n_launch["loss"] - 3 + p.n_launch["bwd"] + (3 if optimizer_step else 0)
