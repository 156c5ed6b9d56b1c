"""Profiling helper (inference path): the dominant conv launches of the batch-B forward pass in isolation - BasicBlock conv2 of
branch 0 / 1 (folded BN, residual, ReLU, lean epilogue), bottleneck conv3, the 480 -> 480 head conv, a fuse-sum host conv - and
the decode / loss kernels at the same batch; each launched twice (warm-up + profiled) with an L2 flush in between:
    ncu --set full -k regex:"conv_tc|softmax_softargmax|decode_argmax|softargmax|final_preds|heatmap_loss" ... python tools/kernel_once_infer.py 256"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrnet_b200 import _lib  # noqa: E402
from hrnet_b200.ops import ConvLayer, PF8, PhasePF8, phase_split  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
lib = _lib.lib()


def conv_case(H, cin, cout, k, res):
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    layer = ConvLayer(w, torch.ones(cout, device=dev), torch.zeros(cout, device=dev), relu=True)
    x = PF8(B, cin, H, H); x.buf.normal_()
    r = PF8(B, cout, H, H) if res else None
    out = PF8(B, cout, H, H)
    return lambda: layer(x, out, r)


cases = [conv_case(64, 32, 32, 3, True), conv_case(32, 64, 64, 3, True), conv_case(16, 128, 128, 3, True),
         conv_case(64, 64, 256, 1, True), conv_case(64, 480, 480, 1, False)]
# fuse-sum host: stride-2 conv from branch 0 producing output 1 with two up-sampled sources (stage 4)
w = torch.randn(64, 32, 3, 3, device=dev) / (32 * 9) ** 0.5
host = ConvLayer(w, torch.ones(64, device=dev), torch.zeros(64, device=dev), stride=2, relu=False)
x0 = PF8(B, 32, 64, 64); x0.buf.normal_()
ph = PhasePF8(B, 32, 64, 64); phase_split(x0, ph)
x1, z2, z3, o1 = PF8(B, 64, 32, 32), PF8(B, 64, 16, 16), PF8(B, 64, 8, 8), PF8(B, 64, 32, 32)
cases.append(lambda: host(ph, o1, x1, fuse=[(z2, 1), (z3, 2)], relu=True))

J, h, wd = 21, 64, 64
logits = torch.randn(B, J, h, wd, device=dev)
heat = torch.softmax(logits.view(B, J, -1), 2).view(B, J, h, wd).contiguous()
gt = torch.rand(B, J, h, wd, device=dev)
out_heat, d_pred = torch.empty_like(heat), torch.empty_like(heat)
coords, maxv = torch.empty(B, J, 2, device=dev), torch.empty(B, J, device=dev)
center, scale = torch.rand(B, 2, device=dev) * 100 + 100, torch.rand(B, 2, device=dev) + 0.5
temp, loss, one, ws = torch.ones(1, device=dev), torch.zeros(1, device=dev), torch.ones(1, device=dev), torch.empty(4096, device=dev)
sp = _lib.stream_ptr
BJ = B * J
cases += [
    lambda: _lib.check(lib.hrnb_decode_argmax(heat.data_ptr(), BJ, h, wd, 0, 1, coords.data_ptr(), maxv.data_ptr(), None, sp())),
    lambda: _lib.check(lib.hrnb_final_preds(heat.data_ptr(), B, J, h, wd, center.data_ptr(), scale.data_ptr(), 1, coords.data_ptr(), maxv.data_ptr(), sp())),
    lambda: _lib.check(lib.hrnb_softargmax(heat.data_ptr(), BJ, h, wd, coords.data_ptr(), sp())),
    lambda: _lib.check(lib.hrnb_softmax_softargmax(logits.data_ptr(), temp.data_ptr(), BJ, h, wd, out_heat.data_ptr(), coords.data_ptr(), sp())),
    lambda: _lib.check(lib.hrnb_loss_heatmap(heat.data_ptr(), gt.data_ptr(), BJ, h * wd, 0, loss.data_ptr(), d_pred.data_ptr(), one.data_ptr(), ws.data_ptr(), sp())),
]
for rep in range(2):
    for fn in cases:
        flush.zero_()
        fn()
torch.cuda.synchronize()
print("done", len(cases), "cases, batch", B)
