"""Flat fp32 parameter / gradient / Adam-moment buffers of the training path.

Every nn.Parameter of the network becomes a view into ONE flat fp32 buffer (so the optimizer is a single kernel and
the data-parallel gradient exchange a single all-reduce); gradients live in a second flat buffer in the layout the
wgrad kernel produces ([tap][cin][cout] for conv weights, natural for 1-D tensors) and are mapped back to the natural
parameter layout only when the nn.Module's `.grad` is asked for.  Optimizer semantics: torch.optim.Adam with L2
weight decay, as `get_optimizer` builds it in the reference (lib/utils/utils.py:71-92, lr 1e-3, wd 1e-4)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ParamSeg

SEG_BLOCK = 1024


class FlatParams:
    def __init__(self, params, conv_meta=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        """params: list of nn.Parameter (CUDA fp32); conv_meta: {index: (cout, cin, cin_g, taps)} for conv weights
        whose gradient is produced in [tap][cin_g][cout] layout (natural layout [cout][cin][taps])."""
        self.params = list(params)
        conv_meta = conv_meta or {}
        dev = self.params[0].device
        self.device = dev
        p_off, g_off, block0 = 0, 0, 0
        segs = (ParamSeg * len(self.params))()
        block_seg = []
        self.p_offs, self.g_offs, self.g_shapes = [], [], []
        for i, p in enumerate(self.params):
            assert p.is_cuda and p.dtype == torch.float32
            n = p.numel()
            s = segs[i]
            s.p_off, s.g_off, s.numel = p_off, g_off, n
            if i in conv_meta:
                cout, cin, cin_g, taps = conv_meta[i]
                assert cout * cin * taps == n, (i, conv_meta[i], n)
                s.cout, s.cin, s.cin_g, s.taps = cout, cin, cin_g, taps
                gn, gshape = taps * cin_g * cout, (taps, cin_g, cout)
            else:
                s.cout = s.cin = s.cin_g = s.taps = 0
                gn, gshape = n, tuple(p.shape)
            s.block0 = block0
            s.frozen = 0 if p.requires_grad else 1
            nb = max(1, (n + SEG_BLOCK - 1) // SEG_BLOCK)
            block_seg.append(np.full(nb, i, dtype=np.int32))
            block0 += nb
            self.p_offs.append(p_off)
            self.g_offs.append(g_off)
            self.g_shapes.append(gshape)
            p_off += (n + 3) // 4 * 4
            g_off += (gn + 3) // 4 * 4
        self.n_params, self.n_grads, self.nblocks = p_off, g_off, block0
        self.data = torch.zeros(p_off, dtype=torch.float32, device=dev)
        self.m = torch.zeros_like(self.data)
        self.v = torch.zeros_like(self.data)
        self.grads = torch.zeros(g_off, dtype=torch.float32, device=dev)
        self._natural = None
        with torch.no_grad():
            for p, off in zip(self.params, self.p_offs):
                view = self.data[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        raw = np.frombuffer(bytes(segs), dtype=np.uint8).copy()
        self.segs_dev = torch.from_numpy(raw).to(dev)
        self.block_seg_dev = torch.from_numpy(np.concatenate(block_seg)).to(dev)
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, weight_decay, 1.0, 1.0, grad_scale], dtype=torch.float32,
                                  device=dev)
        self.step = torch.zeros(1, dtype=torch.int64, device=dev)

    # ---- views ----------------------------------------------------------------------------------------------
    def grad_view(self, i):
        """the gradient segment of parameter i in the layout the kernels write"""
        shape = self.g_shapes[i]
        n = int(np.prod(shape)) if len(shape) else 1
        return self.grads[self.g_offs[i]:self.g_offs[i] + n].view(shape)

    def set_lr(self, lr):
        self.hyper[0] = lr

    def set_grad_scale(self, s):
        self.hyper[7] = s

    def set_grad_from_natural(self, i, g):
        """test helper: store a natural-layout gradient into the kernel layout"""
        gv = self.grad_view(i)
        s = ParamSeg.from_buffer_copy(self.segs_dev[i * C.sizeof(ParamSeg):(i + 1) * C.sizeof(ParamSeg)].cpu().numpy().tobytes())
        if s.taps == 0:
            gv.copy_(g.reshape(gv.shape))
        else:
            gv.zero_()
            gv[:, :s.cin, :].copy_(g.reshape(s.cout, s.cin, s.taps).permute(2, 1, 0))

    # ---- kernels --------------------------------------------------------------------------------------------
    def adam_step(self):
        lib = _lib.lib()
        _lib.check(lib.hrnb_adam_tick(self.hyper.data_ptr(), self.step.data_ptr(), _lib.stream_ptr()))
        _lib.check(lib.hrnb_adam_step(self.data.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.grads.data_ptr(),
                                      self.segs_dev.data_ptr(), self.block_seg_dev.data_ptr(), self.nblocks,
                                      self.hyper.data_ptr(), _lib.stream_ptr()))

    def natural_grads(self, fresh=False):
        """list of gradient tensors in the parameters' own layout: views of one flat buffer that is overwritten by the
        next call, or of a newly allocated one when fresh=True (what autograd's .grad accumulation gets)"""
        if fresh:
            nat = torch.zeros_like(self.data)
        else:
            if self._natural is None:
                self._natural = torch.zeros_like(self.data)
            nat = self._natural
        _lib.check(_lib.lib().hrnb_grad_to_natural(self.grads.data_ptr(), nat.data_ptr(), self.segs_dev.data_ptr(),
                                                   self.block_seg_dev.data_ptr(), self.nblocks, _lib.stream_ptr()))
        return [nat[off:off + p.numel()].view(p.shape) for p, off in zip(self.params, self.p_offs)]

    # ---- checkpointing (tools/train.py:285,375-383: checkpoint['optimizer'] = optimizer.state_dict()) -----------------
    def _trainable(self):
        """indices of the parameters a reference optimizer would hold: filter(requires_grad, model.parameters())
        (lib/utils/utils.py:71-92), in named_parameters() order"""
        return [i for i, p in enumerate(self.params) if p.requires_grad]

    def state_dict(self):
        """torch.optim.Adam-compatible state dict (index-keyed `state`, one param group): a checkpoint written by the fused
        optimizer resumes under the reference's torch.optim.Adam and vice versa."""
        lr, b1, b2, eps, wd = (float(x) for x in self.hyper[:5].cpu())
        step = int(self.step.item())
        idx = self._trainable()
        state = {}
        if step > 0:
            for j, i in enumerate(idx):
                off, n, shape = self.p_offs[i], self.params[i].numel(), self.params[i].shape
                state[j] = {"step": torch.tensor(float(step)),
                            "exp_avg": self.m[off:off + n].view(shape).clone(),
                            "exp_avg_sq": self.v[off:off + n].view(shape).clone()}
        group = {"lr": lr, "betas": (b1, b2), "eps": eps, "weight_decay": wd, "amsgrad": False, "maximize": False,
                 "params": list(range(len(idx)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """inverse of state_dict(); accepts the optimizer state of a reference checkpoint (same parameter order)."""
        groups = sd["param_groups"]
        if len(groups) != 1:
            raise ValueError("the fused Adam holds one parameter group (the reference builds one, lib/utils/utils.py:71-92)")
        g = groups[0]
        idx = self._trainable()
        if len(g["params"]) != len(idx):
            raise ValueError("optimizer state holds %d parameters, the network has %d trainable ones" % (len(g["params"]), len(idx)))
        if g.get("amsgrad", False):
            raise ValueError("amsgrad is not supported by the fused Adam")
        self.hyper[0], self.hyper[1], self.hyper[2] = float(g["lr"]), float(g["betas"][0]), float(g["betas"][1])
        self.hyper[3], self.hyper[4] = float(g["eps"]), float(g["weight_decay"])
        steps = set()
        self.m.zero_()
        self.v.zero_()
        for j, i in enumerate(idx):
            st = sd["state"].get(g["params"][j])
            if st is None:
                continue
            off, n = self.p_offs[i], self.params[i].numel()
            if st["exp_avg"].numel() != n:
                raise ValueError("optimizer state of parameter %d has %d elements, expected %d" % (j, st["exp_avg"].numel(), n))
            self.m[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ (%s): the fused Adam keeps one step counter" % sorted(steps))
        step = steps.pop() if steps else 0
        self.step.fill_(step)
        b1, b2 = float(g["betas"][0]), float(g["betas"][1])
        self.hyper[5], self.hyper[6] = 1.0 - b1 ** step if step else 1.0, 1.0 - b2 ** step if step else 1.0
