"""Profiling helper: the dominant tensor-pipe launches of the training step in isolation (batch 64, HRNet-W32 branch 0
and branch 1 shapes): forward conv, data-gradient conv (accumulating), weight gradient.  Each is launched twice (warm-up +
profiled):  ncu --set full -k regex:"conv_tc|wgrad_tc|bn_stats" --launch-skip ... python tools/kernel_once.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrnet_b200 import tops  # noqa: E402
from hrnet_b200.ops import ConvLayer, PF8  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for H, C_ in ((64, 32), (32, 64)):
    w = torch.randn(C_, C_, 3, 3, device="cuda") / (C_ * 9) ** 0.5
    x = PF8(B, C_, H, H); x.buf.normal_()
    dy = PF8(B, C_, H, H); dy.buf.normal_()
    c = PF8(B, C_, H, H)
    gx = PF8(B, C_, H, H)
    dw = torch.zeros(9, C_, C_, device="cuda")
    sums, sums2 = torch.zeros(C_, 2, device="cuda"), torch.zeros(C_, 2, device="cuda")
    fwd = ConvLayer(w)
    dgrad = ConvLayer(w, transpose=True, tap_ids=[8 - t for t in range(9)])
    for rep in range(2):
        flush.zero_()
        fwd(x, c)                         # forward conv (training: no BN fold, no ReLU)
        flush.zero_()
        fwd(x, c, bn=C_, stats=sums)      # the same conv reducing the BatchNorm batch statistics in its epilogue
        flush.zero_()
        tops.bn_stats(c, sums2)           # the separate statistics pass it replaces
        flush.zero_()
        dgrad(dy, gx, res=gx)             # data gradient accumulated into an existing gradient buffer
        flush.zero_()
        tops.wgrad_conv(dy, x, dw, 3, 1)  # weight gradient
    torch.cuda.synchronize()
print("done")
