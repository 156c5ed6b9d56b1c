"""Achieved HBM bandwidth of the decode / loss kernels at a batch large enough to leave the launch-latency regime
("measure decode at large B (>= 256 => >= 88 MB)").

    python tools/decode_bench.py [B=256] [reps=20] [h=64] [w=64]      -> one JSON line per kernel + a summary line
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none python tools/decode_bench.py 256 2

Each kernel is called through the C ABI on preallocated buffers, the L2 is flushed (a 256 MB fill) before every timed
launch, time = median of `reps` CUDA-event pairs on the launching stream.  Algorithmic bytes: argmax / soft-argmax read the
map once (h*w*4 per map); softmax+soft-argmax reads logits and writes the heat map; HeatmapLoss reads pred + gt (+ writes
d_pred)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrnet_b200 import _lib  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    w = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    J = 21
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    dev = torch.device("cuda")
    lib = _lib.lib()
    g = torch.Generator(device=dev).manual_seed(0)
    logits = torch.randn(B, J, h, w, device=dev, generator=g)
    heat = torch.softmax(logits.view(B, J, -1), 2).view(B, J, h, w).contiguous()
    gt = torch.rand(B, J, h, w, device=dev, generator=g)
    out_heat = torch.empty_like(heat)
    d_pred = torch.empty_like(heat)
    coords = torch.empty(B, J, 2, device=dev)
    maxv = torch.empty(B, J, device=dev)
    center = torch.rand(B, 2, device=dev) * 100 + 100
    scale = torch.rand(B, 2, device=dev) + 0.5
    temp = torch.ones(1, device=dev)
    loss = torch.zeros(1, device=dev)
    ws = torch.empty(4096, device=dev)
    one = torch.ones(1, device=dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    BJ, hw = B * J, h * w
    map_bytes = BJ * hw * 4
    sp = _lib.stream_ptr
    cases = [
        ("decode_argmax (get_max_preds)", map_bytes,
         lambda: lib.hrnb_decode_argmax(heat.data_ptr(), BJ, h, w, 0, 1, coords.data_ptr(), maxv.data_ptr(), None, sp())),
        ("final_preds (argmax + 1/4 px + affine)", map_bytes,
         lambda: lib.hrnb_final_preds(heat.data_ptr(), B, J, h, w, center.data_ptr(), scale.data_ptr(), 1, coords.data_ptr(), maxv.data_ptr(), sp())),
        ("softargmax (get_final_preds use_softmax)", map_bytes,
         lambda: lib.hrnb_softargmax(heat.data_ptr(), BJ, h, w, coords.data_ptr(), sp())),
        ("softmax_softargmax (logits -> heat map + coords)", 2 * map_bytes,
         lambda: lib.hrnb_softmax_softargmax(logits.data_ptr(), temp.data_ptr(), BJ, h, w, out_heat.data_ptr(), coords.data_ptr(), sp())),
        ("softmax_softargmax (coords only)", map_bytes,
         lambda: lib.hrnb_softmax_softargmax(logits.data_ptr(), temp.data_ptr(), BJ, h, w, None, coords.data_ptr(), sp())),
        ("heatmap_loss forward", 2 * map_bytes,
         lambda: lib.hrnb_loss_heatmap(heat.data_ptr(), gt.data_ptr(), BJ, hw, 0, loss.data_ptr(), None, one.data_ptr(), ws.data_ptr(), sp())),
        ("heatmap_loss forward + gradient", 3 * map_bytes,
         lambda: lib.hrnb_loss_heatmap(heat.data_ptr(), gt.data_ptr(), BJ, hw, 0, loss.data_ptr(), d_pred.data_ptr(), one.data_ptr(), ws.data_ptr(), sp())),
    ]
    rows = []
    for name, nbytes, fn in cases:
        _lib.check(fn())
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(fn())
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        med = ts[len(ts) // 2]
        gbs = nbytes / (med * 1e-3) / 1e9
        row = {"kernel": name, "B": B, "J": J, "h": h, "w": w, "algorithmic_bytes": nbytes, "median_us": med * 1e3, "min_us": ts[0] * 1e3,
               "achieved_gbs": gbs, "frac_of_measured_hbm_peak": gbs / peak}
        rows.append(row)
        print(json.dumps(row))
    print(json.dumps({"summary": {r["kernel"]: round(r["frac_of_measured_hbm_peak"], 3) for r in rows}, "peak_gbs": peak, "B": B}))


if __name__ == "__main__":
    main()
