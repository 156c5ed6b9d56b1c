"""Micro-benchmark of hrnb_conv on the HRNet layer shapes (debug/tuning tool, not product).

  python tools/conv_bench.py [--batch 64] [--shapes b0,b1,...] [--reps 20] [--bn N --mb M] [--json out]

Timing: the GPU is first parked behind a long sleep kernel so that every launch and event of the timed
sequence is already queued when it starts (no host-launch gaps inside the event pairs); 4 rotating buffer
sets (> 126 MB in total for the big shapes) defeat L2 residency between repetitions.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {  # name: (hw_out, cin, cout, k, stride, relu, res)   W32 @ 256x256 input
    "b0": (64, 32, 32, 3, 1, True, True), "b1": (32, 64, 64, 3, 1, True, True),
    "b2": (16, 128, 128, 3, 1, True, True), "b3": (8, 256, 256, 3, 1, True, True),
    "l1c1": (64, 256, 64, 1, 1, True, False), "l1c2": (64, 64, 64, 3, 1, True, False),
    "l1c3": (64, 64, 256, 1, 1, True, True), "t1": (64, 256, 32, 3, 1, True, False),
    "conv2": (64, 64, 64, 3, 2, True, False), "t1s2": (32, 256, 64, 3, 2, True, False),
    "head0": (64, 480, 480, 1, 1, True, False), "head3": (64, 480, 21, 1, 1, False, False),
    "f01": (32, 64, 32, 1, 1, False, False), "f02": (16, 128, 32, 1, 1, False, False),
    "f03": (8, 256, 32, 1, 1, False, False), "f10": (32, 32, 64, 3, 2, False, False),
    "f32": (8, 128, 256, 3, 2, False, False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--shapes", default=",".join(SHAPES))
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--mb", type=int, default=0)
    ap.add_argument("--json", default=None)
    ap.add_argument("--dbg", type=int, default=0)
    ap.add_argument("--pdl-off", action="store_true")
    ap.add_argument("--once", action="store_true", help="one launch per shape (for ncu)")
    args = ap.parse_args()
    import torch
    from hrnet_b200.ops import ConvLayer, PF8
    dev = torch.device("cuda")
    from hrnet_b200 import _lib as _l
    _l.lib().hrnb_debug_set(3, args.dbg)
    _l.lib().hrnb_debug_set(2, 1 if args.pdl_off else 0)
    results = []
    for name in args.shapes.split(","):
        hw, cin, cout, k, stride, relu, use_res = SHAPES[name]
        N = args.batch
        nchw = cout % 16 != 0
        w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
        layer = ConvLayer(w, torch.ones(cout, device=dev), torch.zeros(cout, device=dev), stride=stride, relu=relu,
                          out_nchw=nchw)
        nset = 1 if args.once else 4
        sets = []
        for _ in range(nset):
            x = PF8(N, cin, hw * stride, hw * stride)
            x.buf.normal_()          # values irrelevant for timing (padding not zero: results are not checked here)
            out = torch.empty(N, cout, hw, hw, device=dev) if nchw else PF8(N, cout, hw, hw)
            res = PF8(N, cout, hw, hw) if use_res else None
            sets.append((x, out, res))
        prm = [layer.params(x, out, res, mb=args.mb or None, bn=args.bn or None) for x, out, res in sets]
        import ctypes as C
        from hrnet_b200 import _lib
        lib = _lib.lib()

        def launch(i):
            _lib.check(lib.hrnb_conv(C.byref(prm[i % nset]), _lib.stream_ptr()))
        if args.once:
            launch(0)
            torch.cuda.synchronize()
            continue
        for i in range(4):
            launch(i)
        torch.cuda.synchronize()
        torch.cuda._sleep(int(2e7))
        evs = []
        for i in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); launch(i); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        med = ts[len(ts) // 2]
        flops = 2.0 * N * hw * hw * cout * cin * k * k
        byts = 2.0 * N * (cin * (hw * stride) ** 2 + cout * hw * hw * (2 if use_res else 1)) if not nchw else \
            N * (2.0 * cin * hw * hw + 4.0 * cout * hw * hw)
        r = {"dbg": args.dbg, "shape": name, "bn": prm[0].BN, "mb": prm[0].MB, "us": med * 1e3, "min_us": ts[0] * 1e3,
             "tflops": flops / med / 1e9, "gbs": byts / med / 1e6}
        results.append(r)
        print(json.dumps(r), flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
