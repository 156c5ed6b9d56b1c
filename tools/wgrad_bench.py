"""Micro-benchmark (GPU box): wgrad_tc_kernel on the HRNet-W32 layer shapes at batch B over tile candidates
(TG taps per CTA, K splits, KP positions per stage).  python tools/wgrad_bench.py [B]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrnet_b200 import _lib, tops  # noqa: E402
from hrnet_b200.ops import PF8  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
SHAPES = [(64, 32, 32, 3), (32, 64, 64, 3), (16, 128, 128, 3), (8, 256, 256, 3), (64, 64, 256, 1), (64, 256, 64, 1),
          (64, 64, 64, 3), (64, 480, 480, 1), (16, 256, 128, 1), (8, 256, 32, 1)]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")


def timeit(p, reps=5):
    lib = _lib.lib()
    ts = []
    for _ in range(reps):
        flush.zero_()                      # evict L2 between timed launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = lib.hrnb_wgrad(C.byref(p), _lib.stream_ptr())
        b.record()
        b.synchronize()
        if rc != 0:
            return None
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for H, cin, cout, k in SHAPES:
    dy = PF8(B, (cout + 7) // 8 * 8, H, H)
    x = PF8(B, cin, H, H)
    dy.buf.normal_(); x.buf.normal_()
    dw = torch.zeros(k * k, cin, cout, device="cuda")
    taps = tops.fwd_taps_s1(k, dy.Wp)
    rows = []
    base = timeit(tops.wgrad_params(dy, x.ptr, x.ps, dw, cin, cout, taps))
    for TG in ((1, 2, 3, 5, 9) if k == 3 else (1,)):
        for ks in (0, 1, 2, 4, 8, 16, 32, 64):
            for KP in (0, 128):
                t = timeit(tops.wgrad_params(dy, x.ptr, x.ps, dw, cin, cout, taps, TG=TG, ksplit=ks, KP=KP))
                if t is not None:
                    rows.append((t, TG, ks, KP))
    rows.sort()
    print("H%d %d->%d k%d  default %.1f us | best:" % (H, cin, cout, k, base), ["%.1f us TG%d ks%d KP%d" % r for r in rows[:4]], flush=True)
