#!/usr/bin/env python
"""bench.py - HRNet hand-pose hot path throughput on B200 (contract: see DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--width 32|48] [--impl b200|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N

--mode train (default; BASELINE configs[1]): a step = one TRAINING step of pose_hrnet_softmax HRNet-W32 256x256 on
one batch of synthetic images: forward with batch-statistics BatchNorm, spatial softmax + soft-argmax decode,
HeatmapLoss + 0.1 * JointsMSELoss (pose2d), full backward, gradient all-reduce over NCCL when N > 1, fused Adam
(lr 1e-3, L2 wd 1e-4) and weight re-pack.  --mode infer: forward + softmax + soft-argmax decode only.
Weights random-init (reference default init, seed 0), images ~ N(0,1), targets: sigma-2 Gaussians (hrnet_b200.synthetic; the CPU arms use the identical oracle/fixtures).
`value` is images/s with the batch already resident in HBM; `e2e` is the same metric through the public API with
pinned-host inputs copied every step and the losses (train) / decoded joints (infer) read back every step.
`--impl reference` times the CPU restatement of the reference (oracle/, torch-CPU fp32, all host threads) on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_IMG = {(32, 256, 256): 22584492032, (48, 256, 256): 46731362304, (48, 384, 288): 78859173888}
CONV1_FLOP = {256 * 256: 2 * 64 * 27 * 128 * 128}     # stem conv1 has no data-gradient (3 x fwd - conv1 dgrad)
METRIC = "HRNet-W32 256x256 images/sec fwd (forward + softmax soft-argmax decode)"
METRIC_TRAIN = "HRNet-W32 256x256 images/sec fwd+bwd (training step: forward, hm+pose2d loss, backward, Adam)"


def train_flop_per_img(width, H, W):
    f = FLOP_PER_IMG.get((width, H, W))
    if f is None:
        return None
    return 3 * f - 2 * 64 * 27 * (H // 2) * (W // 2)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(width, H, W, device):
    import torch
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet_softmax
    cfg = make_cfg(width, image_size=(H, W))
    torch.manual_seed(0)
    model = pose_hrnet_softmax.get_pose_net(cfg, is_train=False).eval().to(device)
    return model, cfg


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_port_throughput(width, H, W, batch, steps, warmup):
    import numpy as np
    import torch
    from oracle import decode_oracle, fixtures, hrnet_oracle
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet_softmax
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = make_cfg(width, image_size=(H, W))
    torch.manual_seed(0)
    sd = pose_hrnet_softmax.get_pose_net(cfg, is_train=False).state_dict()
    arch = hrnet_oracle.Arch.from_cfg(cfg)
    x = fixtures.images(batch, H, W)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        heat = hrnet_oracle.forward(sd, x, arch, "softmax")[0]
        decode_oracle.spatial_expectation2d(heat.numpy())
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, total / len(times) * 1e3, cores


def cpu_port_train_throughput(width, H, W, batch, steps, warmup):
    """the reference's training step restated on the CPU (oracle/train_oracle.py): forward (train-mode BN), losses,
    autograd backward, torch.optim.Adam step"""
    import torch
    from oracle import fixtures, hrnet_oracle, train_oracle
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet_softmax
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = make_cfg(width, image_size=(H, W))
    torch.manual_seed(0)
    sd = pose_hrnet_softmax.get_pose_net(cfg, is_train=False).state_dict()
    arch = hrnet_oracle.Arch.from_cfg(cfg)
    x = fixtures.images(batch, H, W)
    gt, xy, vis = fixtures.targets(batch, 21, H // 4, W // 4)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        train_oracle.train_step(sd, x, gt, xy, vis, arch, "softmax")
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, total / len(times) * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bs = min(args.batch, 8)
    train = args.mode == "train"
    steps, warm = (min(args.steps, 4), min(args.warmup, 1)) if train else (args.steps, args.warmup)
    fn = cpu_port_train_throughput if train else cpu_port_throughput
    ips, ms, cores = fn(args.width, args.height, args.img_width, bs, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC_TRAIN if train else METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, bs, note="CPU restatement of the reference (oracle port, torch-CPU fp32), "
                                  "bounded sample: batch %d per step" % bs),
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x batch %d of the same workload (after %d warm-up)" % (steps, bs, warm)},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, batch, note=None):
    if args.mode == "train":
        wl = ("pose_hrnet_softmax HRNet-W%d %dx%d TRAINING step (BASELINE configs[1]): forward with batch-stat BN, softmax + "
              "soft-argmax, HeatmapLoss + 0.1*pose2d loss, backward, %sfused Adam, weight re-pack; 21 joints, batch %d/GPU"
              % (args.width, args.height, args.img_width, "NCCL gradient all-reduce, " if args.gpus > 1 else "", batch))
    else:
        wl = ("pose_hrnet_softmax HRNet-W%d %dx%d inference forward + spatial softmax + soft-argmax decode, "
              "21 joints, batch %d/GPU (BASELINE configs[1] geometry)" % (args.width, args.height, args.img_width, batch))
    c = {"workload": wl,
         "batch_per_gpu": batch, "global_batch": batch * args.gpus, "parallelism": ("dp%d (batch sharded, one fp32 gradient all-reduce per step)" if args.mode == "train" else "dp%d (batch sharded, no collective)") % args.gpus,
         "l2": "each step streams ~%.1f GB of activations (>> 126 MB L2) and reads a fresh input batch from a pool of 4 "
               "(4 x %.0f MB > L2); no explicit flush" % (0.061 * batch * (args.height * args.img_width) / 65536.0,
                                                         batch * 3 * args.height * args.img_width * 4 / 1e6)}
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def per_kernel_conv_timing(plan, torch, reps=3):
    """Eager, single-stream passes, every launch queued behind a parked GPU so that host launch gaps do not count:
    (1) one CUDA-event pair around the whole pass -> serial step time; (2) an event pair around every launch ->
    each kernel's SHARE of the step (the pairs themselves add ~2 us per launch, so only the shares are used).
    Time of the dominant kernel (conv_tc_kernel, all launches) = serial step time x its share.
    Returns (conv_ms, serial_ms, n_conv, per-launch list)."""
    steps = [s for s in plan.steps if s.kind == "op"]
    best_tot, best = None, None
    for _ in range(reps):
        torch.cuda.synchronize()
        torch.cuda._sleep(int(4e7))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in steps:
            s.fn()
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b)
        best_tot = t if best_tot is None else min(best_tot, t)
        evs = []
        torch.cuda._sleep(int(4e7))
        for s in steps:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s.fn(); b.record()
            evs.append((s.name, a, b))
        torch.cuda.synchronize()
        d = [(n, a.elapsed_time(b)) for n, a, b in evs]
        tot = sum(t for _, t in d)
        if best is None or tot < best[0]:
            best = (tot, d)
    d = best[1]
    is_conv = lambda n: (n not in ("conv1.im2col", "softmax_softargmax", "decode_argmax") and ".fuse." not in n
                         and "bilinear" not in n and "split" not in n)
    share = sum(t for n, t in d if is_conv(n)) / best[0]
    return best_tot * share, best_tot, sum(1 for n, _ in d if is_conv(n)), d


def run_b200(args):
    if args.mode == "train":
        return run_b200_train(args)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from hrnet_b200 import _lib
    from hrnet_b200 import synthetic as fixtures     # the product arm never imports oracle/
    H, W, B = args.height, args.img_width, args.batch
    model, cfg = build_model(args.width, H, W, dev)
    model.return_features = False      # the decode path does not consume the 480-channel feature tensor
    model.static_outputs = True
    eng = model.engine()
    plan = eng.plan(B, H, W)
    pool = [fixtures.images(B, H, W, seed=1 + rank * 16 + i).to(dev) for i in range(4)]
    hbm, tf_burst, tf_sust, peak_kind = peaks()
    flop_img = FLOP_PER_IMG.get((args.width, H, W))

    def step(i):
        plan.x.copy_(pool[i % 4], non_blocking=True)
        plan.run(want_features=False, use_graph=True)

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    from hrnet_b200.parallel import max_over_ranks
    ms = max_over_ranks(ms, device=dev)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- e2e: public API (pipeline.StreamingPredictor), pinned host input, decoded joints read back each step ----
    from hrnet_b200.pipeline import StreamingPredictor
    host = [fixtures.images(B, H, W, seed=100 + rank * 16 + i).pin_memory() for i in range(3)]
    pred = StreamingPredictor(model)

    def host_batches(n):
        for i in range(n):
            yield host[i % 3]

    for _ in pred.run(host_batches(3)):
        pass
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    n_out = 0
    for joints in pred.run(host_batches(args.steps)):
        n_out += joints.shape[0]          # the user consumes the joints of every step on the host
    g1.record()
    torch.cuda.synchronize()
    assert n_out == B * args.steps
    e2e_ms = max_over_ranks(g0.elapsed_time(g1), device=dev)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0) ------------------------------------------------------
    conv_ms, all_ms, n_conv, detail = per_kernel_conv_timing(plan, torch)
    roof = None
    if flop_img:
        conv_flop_step = flop_img * B       # every conv of the net (incl. the stem) runs in conv_tc_kernel
        achieved = conv_flop_step / (conv_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (all %d launches of one step)" % n_conv,
                "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst,
                "traffic": None, "peak_source": "%s bf16_tflops (burst); duration = eager single-stream step time (one CUDA-event pair, launches "
                "pre-queued) x the conv kernels' share from per-launch event pairs" % peak_kind,
                "avg_launch_us": conv_ms / n_conv * 1e3, "serial_step_ms": all_ms,
                "conv_share_of_serial_step": conv_ms / all_ms,
                "step_frac_of_sustained_peak": (value / world) * flop_img / 1e12 / tf_sust}
    # ---- CPU baseline (bounded sample) ---------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ips, cms, cores = cpu_port_throughput(args.width, H, W, 8, 3, 1)
        cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "oracle port (torch-CPU fp32 + numpy decode), 3 steps x batch 8 after 1 warm-up, %.0f ms/step" % cms}
    line = {
        "metric": METRIC if args.width == 32 else METRIC.replace("W32 256x256", "W%d %dx%d" % (args.width, H, W)),
        "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, B),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * H * W * 4,
                "d2h_bytes_per_step": B * 21 * 2 * 4, "ms_per_step": e2e_ms / args.steps,
                "api": "pipeline.StreamingPredictor(model).run(pinned host batches): H2D copy + model() + get_final_preds + D2H of every batch, copies overlapped with the previous/next batch"},
        "gpu_launches": args.steps * plan.launches(False),
        "roofline": roof, "cpu_baseline": cpu,
        "tensor_frac_of_burst_peak": (value / world) * flop_img / 1e12 / tf_burst if flop_img else None,
    }
    print(json.dumps(line))
    if args.detail:
        with open(args.detail, "w") as f:
            json.dump({"per_launch_ms": detail, "conv_ms": conv_ms, "all_ms": all_ms}, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def timed_kinds(fns, names, torch, reps=2):
    """eager, single stream, launches pre-queued behind a parked GPU: (serial time of the whole list [one event pair],
    {kind: share of the summed per-launch times, count})"""
    best_tot, best = None, None
    for _ in range(reps):
        torch.cuda.synchronize()
        torch.cuda._sleep(int(8e7))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for f in fns:
            f()
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b)
        best_tot = t if best_tot is None else min(best_tot, t)
        evs = []
        torch.cuda._sleep(int(8e7))
        for f in fns:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        d = [a.elapsed_time(b) for a, b in evs]
        if best is None or sum(d) < sum(best):
            best = d
    kinds = {}
    tot = sum(best)
    for n, t in zip(names, best):
        k = n.split(":", 1)[0] if ":" in n else "other"
        e = kinds.setdefault(k, [0.0, 0])
        e[0] += t / tot
        e[1] += 1
    return best_tot, kinds, list(zip(names, best))


def _mark(msg):
    """progress marker on stderr (stdout carries exactly one JSON line)"""
    print("[bench] %s t=%.1fs" % (msg, time.time() - _T0), file=sys.stderr, flush=True)


_T0 = time.time()


def run_b200_train(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from hrnet_b200 import _lib
    from hrnet_b200.parallel import max_over_ranks
    from hrnet_b200.train import TrainEngine
    from hrnet_b200 import synthetic as fixtures     # the product arm never imports oracle/
    H, W, B = args.height, args.img_width, args.batch
    model, cfg = build_model(args.width, H, W, dev)
    model.train()
    eng = TrainEngine(model, lr=1e-3, weight_decay=1e-4, loss_factors=(1.0, 0.1))
    if world > 1:     # same initial weights on every rank (seeded identically) - assert instead of broadcasting blindly
        chk = eng.flat.data.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float(lo) == float(hi), "ranks start from different weights"
    from hrnet_b200.parallel import GradAllReduce
    allreduce = GradAllReduce(eng.flat.grads.numel(), n_buckets=1) if world > 1 else None
    if world > 1:
        eng.flat.set_grad_scale(allreduce.mean_scale)
    plan = eng.plan(B, H, W)
    pool = []
    for i in range(4):
        gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4, seed=2 + rank * 16 + i)
        pool.append((fixtures.images(B, H, W, seed=1 + rank * 16 + i).to(dev), gt.to(dev), xy.to(dev), vis.to(dev)))
    hbm, tf_burst, tf_sust, peak_kind = peaks()
    flop_img = train_flop_per_img(args.width, H, W)

    def step(i):
        x, gt, xy, vis = pool[i % 4]
        eng.train_step(x, gt, xy, vis, allreduce=allreduce)

    _mark("engine + plan built")
    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    _mark("warm-up done")
    loss0 = float(plan.losses[0])
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), device=dev)
    launches = args.steps * eng.launches_per_step(plan)     # graph replays bypass the library's launch counter
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)
    loss1 = float(plan.losses[0])
    _mark("timed region done (%.2f ms/step)" % (ms / args.steps))

    # ---- e2e: TrainEngine.train_step fed from pinned host memory, losses read back to the host every step ----
    host = []
    for i in range(3):
        gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4, seed=200 + rank * 16 + i)
        host.append(tuple(t.pin_memory() for t in (fixtures.images(B, H, W, seed=100 + rank * 16 + i), gt, xy, vis)))
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    def e2e_step(i):
        x, gt, xy, vis = host[i % 3]
        p = eng.train_step(x, gt, xy, vis, allreduce=allreduce)
        return p.losses.cpu()          # device -> host read of [total, heat-map, pose2d] (synchronises the step)

    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(args.steps):
        l = e2e_step(i)
    g1.record()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(g0.elapsed_time(g1), device=dev)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    _mark("e2e done")

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernels: every tensor-pipe launch of one step (conv fwd, dgrad, wgrad) ----
    fns = plan.fwd_fns + plan.loss_steps + plan.bwd_fns
    names = plan.fwd_names + ["loss"] * len(plan.loss_steps) + plan.bwd_names
    serial_ms, kinds, detail = timed_kinds(fns, names, torch)
    _mark("per-kernel timing done")
    tshare = sum(kinds.get(k, [0.0, 0])[0] for k in ("conv", "dgrad", "wgrad"))
    n_tensor = sum(kinds.get(k, [0.0, 0])[1] for k in ("conv", "dgrad", "wgrad"))
    roof = None
    if flop_img:
        tensor_ms = serial_ms * tshare
        achieved = flop_img * B / (tensor_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (forward + data-gradient) and wgrad_tc_kernel: all %d tensor-pipe launches of one step" % n_tensor,
                "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst, "traffic": None,
                "peak_source": "%s bf16_tflops (burst); duration = eager single-stream time of fwd+loss+bwd (one CUDA-event pair, "
                               "launches pre-queued) x the tensor-pipe kernels' share from per-launch event pairs" % peak_kind,
                "avg_launch_us": tensor_ms / max(1, n_tensor) * 1e3, "serial_step_ms": serial_ms,
                "share_of_serial_step": {k: round(v[0], 4) for k, v in sorted(kinds.items())},
                "launches_by_kind": {k: v[1] for k, v in sorted(kinds.items())},
                "step_frac_of_sustained_peak": (value / world) * flop_img / 1e12 / tf_sust}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ips, cms, cores = cpu_port_train_throughput(args.width, H, W, 8, 2, 1)
        cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "oracle port of the training step (torch-CPU fp32 autograd + Adam), 2 steps x batch 8 after 1 warm-up, %.0f ms/step" % cms}
    line = {
        "metric": METRIC_TRAIN if args.width == 32 else METRIC_TRAIN.replace("W32 256x256", "W%d %dx%d" % (args.width, H, W)),
        "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, B),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                "ms_per_step": e2e_ms / args.steps,
                "api": "TrainEngine.train_step(pinned host images, heat-map targets, joints, visibility) + losses.cpu() every step"},
        "gpu_launches": launches,
        "roofline": roof, "cpu_baseline": cpu,
        "tensor_frac_of_burst_peak": (value / world) * flop_img / 1e12 / tf_burst if flop_img else None,
        "loss_first_last": [loss0, loss1], "activation_bytes": plan.act_bytes,
    }
    print(json.dumps(line))
    if args.detail:
        with open(args.detail, "w") as f:
            json.dump({"per_launch_ms": detail, "serial_ms": serial_ms, "kinds": kinds}, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    if "HRNB_KEEP_NCCL_DEBUG" not in os.environ:
        os.environ["NCCL_DEBUG"] = "WARN"       # NCCL's version banner goes to stdout, which must carry exactly one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--width", type=int, default=32)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--img-width", type=int, default=256)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--detail", default=None, help="write per-launch timings (json) here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        try:
            run_b200(args)
        except Exception:
            # a tcgen05 kernel that trapped on its bounded mbarrier wait leaves (kernel, CTA, warp, barrier) records
            try:
                from hrnet_b200 import _lib
                recs = _lib.hang_report()
                if recs:
                    print("[bench] mbarrier time-out records (%d): %s" % (len(recs), recs[:48]), file=sys.stderr)
            except Exception:
                pass
            raise


if __name__ == "__main__":
    main()
