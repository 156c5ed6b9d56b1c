#!/usr/bin/env python
"""bench.py - HRNet hand-pose hot path throughput on B200 (contract: see DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference] [--config 1..5]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N

Default (BASELINE configs[1], the configuration BASELINE.json's metric is quoted on): a step = one TRAINING step of
pose_hrnet_softmax HRNet-W32 256x256 on one batch of 64 synthetic images per GPU: forward with batch-statistics
BatchNorm, spatial softmax + soft-argmax decode, HeatmapLoss + 0.1 * JointsMSELoss (pose2d), full backward, bucketed
gradient all-reduce over NCCL when N > 1 (overlapped with the rest of the backward pass), fused Adam (lr 1e-3, L2 wd
1e-4) and weight re-pack.  The same JSON line carries the FORWARD half of BASELINE's metric ("images/sec fwd /
fwd+bwd") as `infer`: forward + decode at batch 64 and 256 per GPU, each with its own e2e and roofline.

  --mode infer                         forward + decode only (headline = the forward number)
  --config 1|2|3|4|5                   BASELINE.json configs[i-1]: 1 = W32 raw variant + get_max_preds argmax, batch 8,
                                       inference; 2 = the default; 3 = W48 384x288 trainable-softmax inference (use
                                       --sweep 1,2,...,256 for the batch sweep); 4 = MHP 4-view W32 + algebraic DLT
                                       triangulation (inference, batch = samples, 4 views each); 5 = W48 256x256 raw variant,
                                       HeatmapLoss only, training
  --variant softmax|raw  --loss hm+pose2d|hm  --width 32|48  --height H --img-width W  --batch B   (explicit form)

Weights random-init (reference default init, seed 0), images ~ N(0,1), targets: sigma-2 Gaussians (hrnet_b200.synthetic;
the CPU arms use the identical oracle/fixtures).  `value` is images/s with the batch already resident in HBM; `e2e` is
the same metric through the public API (pipeline.StreamingTrainer / StreamingPredictor) with pinned-host inputs copied
every step and the losses (train) / decoded joints (infer) read back every step.
`--impl reference` times the reference's own CPU implementation (the UNMODIFIED reference modules from oracle/_ref when
oracle/build_ref.py has placed them, else the oracle port; all host threads) on a bounded sample (batch 8 per step) of
the same workload, for exactly the --steps / --warmup it is given, and reports the median step.
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
CPU_SAMPLE_BATCH = 8


def flop_per_img(width, H, W):
    """forward conv FLOP per image (reference MAC convention lib/utils/utils.py:154-159, FLOP = 2 MAC); the network is
    fully convolutional, so other input sizes scale with the pixel count of the measured 256x256 figure"""
    f = FLOP_PER_IMG.get((width, H, W))
    if f is None and (width, 256, 256) in FLOP_PER_IMG:
        f = FLOP_PER_IMG[(width, 256, 256)] * (H * W) / 65536.0
    return f


def train_flop_per_img(width, H, W):
    f = flop_per_img(width, H, W)
    if f is None:
        return None
    return 3 * f - 2 * 64 * 27 * (H // 2) * (W // 2)      # stem conv1 has no data-gradient


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


def ncu_traffic(kind):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/r2_ncu_traffic.json,
    written by tools/ncu_traffic.py from the .ncu-rep of the same workload); None when no capture is committed"""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
            return json.load(f).get(kind)
    except Exception:
        return None


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


def build_model(args, device, train=False):
    import torch
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet, pose_hrnet_softmax
    cfg = make_cfg(args.width, softmax=(args.variant == "softmax"), trainable_softmax=args.trainable_temp,
                   image_size=(args.height, args.img_width))
    torch.manual_seed(0)
    mod = pose_hrnet_softmax if args.variant == "softmax" else pose_hrnet
    model = mod.get_pose_net(cfg, is_train=False).to(device)
    return (model.train() if train else model.eval()), cfg


def _median(v):
    v = sorted(v)
    n = len(v)
    return v[n // 2] if n % 2 else 0.5 * (v[n // 2 - 1] + v[n // 2])


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------------
def _reference_modules():
    """the UNMODIFIED reference modules from oracle/_ref (or /root/reference), or None -> oracle port"""
    try:
        from oracle import ref_shim
        if ref_shim.available():
            return ref_shim, ref_shim.modules()
    except Exception as e:       # a broken vendored copy must not take the arm down: fall back to the port, say so
        print("[bench] reference modules unavailable (%s): timing the oracle port" % e, file=sys.stderr)
    return None, None


def _ref_cfg(ref_shim, args):
    """the reference experiment YAML of this workload (w48: NUM_CHANNELS widened as the w48 YAMLs do)"""
    rel = ("experiments/RHD/RHD_HRNet_w32_softmax_hm-pose2dloss_v1.yaml" if args.variant == "softmax"
           else "experiments/RHD/RHD_HRNet_w32_max_hmloss_v1.yaml")
    cfg = ref_shim.load_cfg(rel)
    for s_, nb in ((2, 2), (3, 3), (4, 4)):
        cfg.MODEL.EXTRA["STAGE%d" % s_]["NUM_CHANNELS"] = [args.width * 2 ** i for i in range(nb)]
    if args.variant == "softmax":
        cfg.MODEL["TRAINABLE_SOFTMAX"] = bool(args.trainable_temp)
    return cfg


def cpu_throughput(args, train, batch, steps, warmup):
    """-> (images/s from the MEDIAN step, median ms/step, cores, kind, steps actually timed, warm-up actually run)"""
    import torch
    from oracle import decode_oracle, fixtures, hrnet_oracle, train_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H, W = args.height, args.img_width
    x = fixtures.images(batch, H, W)
    gt, xy, vis = fixtures.targets(batch, 21, H // 4, W // 4)
    ref_shim, mods = _reference_modules()
    softmax = args.variant == "softmax"
    if mods is not None:
        pose_hrnet, pose_hrnet_softmax, hd, inf, loss = mods
        cfg = _ref_cfg(ref_shim, args)
        torch.manual_seed(0)
        model = (pose_hrnet_softmax if softmax else pose_hrnet).get_pose_net(cfg, is_train=False)
        kind = "reference"
        if train:
            model.train()
            opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-3, weight_decay=1e-4)
            hm_loss, p2d_loss = loss.HeatmapLoss(), loss.JointsMSELoss()

            def step():
                out = model(x)
                total = 1.0 * hm_loss(out[0], gt)
                if softmax and args.loss == "hm+pose2d":
                    total = total + 0.1 * p2d_loss(hd.get_final_preds(out[0], True), xy, vis)
                opt.zero_grad()
                total.backward()
                opt.step()
                return float(total.detach())
        else:
            model.eval()

            def step():
                with torch.no_grad():
                    out = model(x)
                if softmax:
                    return hd.get_final_preds(out[0], True)
                return inf.get_max_preds(out[0].numpy())
    else:
        from hrnet_b200.config import make_cfg
        from hrnet_b200.models import pose_hrnet as ph, pose_hrnet_softmax as phs
        cfg = make_cfg(args.width, softmax=softmax, trainable_softmax=args.trainable_temp, image_size=(H, W))
        torch.manual_seed(0)
        sd = (phs if softmax else ph).get_pose_net(cfg, is_train=False).state_dict()
        arch = hrnet_oracle.Arch.from_cfg(cfg)
        kind = "port"
        if train:
            state = {"sd": sd, "opt": None}

            def step():
                o = train_oracle.train_step(state["sd"], x, gt, xy, vis, arch, args.variant, trainable_temp=args.trainable_temp,
                                            f_p2d=(0.1 if args.loss == "hm+pose2d" else 0.0), opt_state=state["opt"])
                state["sd"], state["opt"] = o["state"], o["opt_state"]
        else:
            def step():
                out = hrnet_oracle.forward(sd, x, arch, args.variant)[0]
                if softmax:
                    return decode_oracle.spatial_expectation2d(out.numpy())
                return decode_oracle.get_max_preds(out.numpy())
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = _median(times)
    return batch / med, med * 1e3, cores, kind, len(times), warmup


def cpu_baseline_record(args, train, steps=5, warmup=2):
    ips, ms, cores, kind, n, w = cpu_throughput(args, train, CPU_SAMPLE_BATCH, steps, warmup)
    what = ("UNMODIFIED reference modules (oracle/_ref), torch-CPU fp32" if kind == "reference"
            else "oracle port of the reference (torch-CPU fp32)")
    return {"value": ips, "unit": "images/s", "cores": cores, "kind": kind,
            "sample": "%s: %s, batch %d per step, median of %d steps after %d warm-up (%.0f ms/step)"
                      % ("training step (forward, losses, backward, Adam)" if train else "forward + decode", what,
                         CPU_SAMPLE_BATCH, n, w, ms)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    train = args.mode == "train"
    steps, warm = max(1, args.steps), max(0, args.warmup)
    ips, ms, cores, kind, n, w = cpu_throughput(args, train, CPU_SAMPLE_BATCH, steps, warm)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": n, "warmup": w, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, CPU_SAMPLE_BATCH,
                                  note="the reference's own CPU path (%s, torch-CPU fp32, %d host threads, rank 0 only) on a bounded "
                                       "sample of the workload: batch %d per step instead of %d; value = batch / MEDIAN step time"
                                       % ("UNMODIFIED reference modules from oracle/_ref" if kind == "reference" else "oracle port",
                                          cores, CPU_SAMPLE_BATCH, args.batch)),
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": "%d timed steps x batch %d after %d warm-up, median step %.0f ms" % (n, CPU_SAMPLE_BATCH, w, ms)},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def metric_name(args):
    geo = "HRNet-W%d %dx%d" % (args.width, args.height, args.img_width)
    if args.mode == "train":
        return geo + " images/sec fwd+bwd (training step: forward, %s loss, backward, Adam)" % (
            "hm+pose2d" if (args.variant == "softmax" and args.loss == "hm+pose2d") else "heat-map")
    if args.mode == "mhp":
        return geo + " multi-view samples/sec (4 views: forward + soft-argmax + algebraic DLT triangulation)"
    return geo + " images/sec fwd (forward + %s decode)" % ("softmax soft-argmax" if args.variant == "softmax" else "get_max_preds argmax")


def workload_config(args, batch, note=None):
    var = "pose_hrnet_softmax" if args.variant == "softmax" else "pose_hrnet"
    if args.mode == "train":
        wl = ("%s HRNet-W%d %dx%d TRAINING step (BASELINE configs[%d]): forward with batch-stat BN, %s, backward, %sfused Adam, "
              "weight re-pack; 21 joints, batch %d/GPU"
              % (var, args.width, args.height, args.img_width, args.config - 1 if args.config else 1,
                 "softmax + soft-argmax, HeatmapLoss + 0.1*pose2d loss" if (args.variant == "softmax" and args.loss == "hm+pose2d")
                 else "HeatmapLoss", "bucketed NCCL gradient all-reduce overlapped with the backward pass, " if args.gpus > 1 else "", batch))
        par = "dp%d (batch sharded, fp32 gradient sum-all-reduce in 3 buckets per step, overlapped with the backward pass)" % args.gpus
    elif args.mode == "mhp":
        wl = ("MHP multi-view %s HRNet-W%d %dx%d, 4 views per sample: forward + soft-argmax + algebraic (SII-DLT) triangulation "
              "(BASELINE configs[3]), %d samples (%d images) per GPU" % (var, args.width, args.height, args.img_width, batch, 4 * batch))
        par = "dp%d (sharded by sample, all views of a sample on one rank, no collective)" % args.gpus
    else:
        wl = ("%s HRNet-W%d %dx%d inference forward + %s decode, 21 joints, batch %d/GPU"
              % (var, args.width, args.height, args.img_width,
                 "spatial softmax + soft-argmax" if args.variant == "softmax" else "get_max_preds argmax", batch))
        par = "dp%d (batch sharded, no collective)" % args.gpus
    c = {"workload": wl, "batch_per_gpu": batch, "global_batch": batch * args.gpus, "parallelism": par,
         "l2": "each step streams ~%.1f GB of activations (>> 126 MB L2) and reads a fresh input batch from a pool of 4 "
               "(4 x %.0f MB); no explicit flush" % ((0.061 if args.mode == "train" else 0.012) * batch * (args.height * args.img_width) / 65536.0
                                                     * (args.width / 32.0), batch * 3 * args.height * args.img_width * 4 / 1e6)}
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


def measure_infer(args, B, dev, rank, world, steps, warmup, with_roofline=True):
    """forward + decode at batch B per GPU -> dict(value, ms_per_step, e2e, roofline, gpu_launches, ...) ; every rank
    calls it, the timing is the max over ranks"""
    import torch
    import torch.distributed as dist
    from hrnet_b200 import synthetic as fixtures     # the product arm never imports oracle/
    from hrnet_b200.parallel import max_over_ranks
    from hrnet_b200.pipeline import StreamingPredictor
    H, W = args.height, args.img_width
    model, cfg = build_model(args, dev)
    model.return_features = False      # the decode path does not consume the feature tensor
    model.static_outputs = True
    eng = model.engine()
    plan = eng.plan(B, H, W)
    pool = [fixtures.images(B, H, W, seed=1 + rank * 16 + i).to(dev) for i in range(4)]
    hbm, tf_burst, tf_sust, peak_kind = peaks()
    fimg = flop_per_img(args.width, H, W)

    def step(i):
        plan.x.copy_(pool[i % 4], non_blocking=True)
        plan.run(want_features=False, use_graph=True)

    for i in range(max(warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), device=dev)
    if world > 1:
        dist.barrier()
    value = world * B * steps / (ms / 1e3)

    # ---- e2e: public API (pipeline.StreamingPredictor), pinned host input, decoded joints read back each step ----
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
    for joints in pred.run(host_batches(steps)):
        n_out += joints.shape[0]          # the user consumes the joints of every step on the host
    g1.record()
    torch.cuda.synchronize()
    assert n_out == B * steps
    e2e_ms = max_over_ranks(g0.elapsed_time(g1), device=dev)
    e2e_value = world * B * steps / (e2e_ms / 1e3)
    res = {"batch_per_gpu": B, "value": value, "unit": "images/s", "ms_per_step": ms / steps, "steps": steps,
           "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * H * W * 4,
                   "d2h_bytes_per_step": B * 21 * 2 * 4, "ms_per_step": e2e_ms / steps,
                   "api": "pipeline.StreamingPredictor(model).run(pinned host batches): H2D copy + model() + get_final_preds + D2H "
                          "of every batch, copies overlapped with the previous/next batch"},
           "gpu_launches": steps * plan.launches(False), "launches_per_step": plan.launches(False),
           "tensor_frac_of_burst_peak": (value / world) * fimg / 1e12 / tf_burst if fimg else None,
           "tensor_frac_of_sustained_peak": (value / world) * fimg / 1e12 / tf_sust if fimg else None}
    detail = None
    if with_roofline and rank == 0 and fimg:
        conv_ms, all_ms, n_conv, detail = per_kernel_conv_timing(plan, torch)
        achieved = fimg * B / (conv_ms / 1e3) / 1e12
        res["roofline"] = {"bound": "tensor", "kernel": "conv_tc_kernel (all %d launches of one step)" % n_conv,
                           "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst,
                           "traffic": ncu_traffic("conv_tc_kernel_infer"),
                           "peak_source": "%s bf16_tflops (burst); duration = eager single-stream step time (one CUDA-event pair, "
                                          "launches pre-queued) x the conv kernels' share from per-launch event pairs" % peak_kind,
                           "avg_launch_us": conv_ms / n_conv * 1e3, "serial_step_ms": all_ms,
                           "conv_share_of_serial_step": conv_ms / all_ms}
    del plan, eng, model, pool, pred, host
    torch.cuda.empty_cache()
    return res, detail


def run_b200_infer(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world, dev = _dist_setup()
    batches = [int(b) for b in args.sweep.split(",")] if args.sweep else [args.batch]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    results, detail = [], None
    for B in batches:
        r, d = measure_infer(args, B, dev, rank, world, args.steps, args.warmup)
        results.append(r)
        detail = detail or d
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        head = results[-1] if args.sweep else results[0]
        cpu = cpu_baseline_record(args, False) if (world == 1 and not args.no_cpu_baseline) else None
        line = {"metric": metric_name(args), "value": head["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, head["batch_per_gpu"]),
                "clocks": clocks, "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head.get("roofline"),
                "cpu_baseline": cpu, "tensor_frac_of_burst_peak": head["tensor_frac_of_burst_peak"]}
        if args.sweep:
            line["sweep"] = results
        print(json.dumps(line))
        if args.detail and detail:
            with open(args.detail, "w") as f:
                json.dump({"per_launch_ms": detail}, f, indent=1)
    _dist_teardown(world)


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


def _dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return rank, local_rank, world, dev


def _dist_teardown(world):
    import torch.distributed as dist
    if world > 1 and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def run_b200_train(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world, dev = _dist_setup()
    from hrnet_b200.parallel import GradAllReduce, max_over_ranks
    from hrnet_b200.pipeline import StreamingTrainer
    from hrnet_b200.train import TrainEngine
    from hrnet_b200 import synthetic as fixtures     # the product arm never imports oracle/
    H, W, B = args.height, args.img_width, args.batch
    model, cfg = build_model(args, dev, train=True)
    f_p2d = 0.1 if (args.variant == "softmax" and args.loss == "hm+pose2d") else 0.0
    eng = TrainEngine(model, lr=1e-3, weight_decay=1e-4, loss_factors=(1.0, f_p2d))
    allreduce = None
    if world > 1:     # same initial weights on every rank (seeded identically) - assert instead of broadcasting blindly
        chk = eng.flat.data.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float(lo) == float(hi), "ranks start from different weights"
        # HRNB_AR_OVERLAP=0: one collective over the whole buffer after the backward graph (the round-1 form), for A/B runs
        allreduce = GradAllReduce(eng.flat.grads.numel(), device=dev if os.environ.get("HRNB_AR_OVERLAP", "1") != "0" else None)
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

    # ---- e2e: pipeline.StreamingTrainer fed from pinned host memory, the losses of every step read back to the host ----
    host = []
    for i in range(3):
        gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4, seed=200 + rank * 16 + i)
        host.append(tuple(t.pin_memory() for t in (fixtures.images(B, H, W, seed=100 + rank * 16 + i), gt, xy, vis)))
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    trainer = StreamingTrainer(eng, allreduce=allreduce)

    def host_batches(n):
        for i in range(n):
            yield host[i % 3]

    for _ in trainer.run(host_batches(3)):
        pass
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    n_loss = 0
    for l in trainer.run(host_batches(args.steps)):
        n_loss += 1 if bool(l[0] == l[0]) else 0         # the user reads the (finite) losses of every step on the host
    g1.record()
    torch.cuda.synchronize()
    assert n_loss == args.steps, (n_loss, args.steps)
    e2e_ms = max_over_ranks(g0.elapsed_time(g1), device=dev)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    _mark("e2e done")

    line = None
    if rank == 0:
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
                    "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst,
                    "traffic": ncu_traffic("conv_tc_kernel_train"),
                    "peak_source": "%s bf16_tflops (burst); duration = eager single-stream time of fwd+loss+bwd (one CUDA-event pair, "
                                   "launches pre-queued) x the tensor-pipe kernels' share from per-launch event pairs" % peak_kind,
                    "avg_launch_us": tensor_ms / max(1, n_tensor) * 1e3, "serial_step_ms": serial_ms,
                    "share_of_serial_step": {k: round(v[0], 4) for k, v in sorted(kinds.items())},
                    "launches_by_kind": {k: v[1] for k, v in sorted(kinds.items())},
                    "step_frac_of_sustained_peak": (value / world) * flop_img / 1e12 / tf_sust}
        line = {
            "metric": metric_name(args), "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "pipeline.StreamingTrainer(engine).run(pinned host batches): per step H2D of images, heat-map targets, "
                           "joints, visibility (copy stream, double-buffered) + TrainEngine.train_step + D2H of the 3 losses"},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": None,
            "tensor_frac_of_burst_peak": (value / world) * flop_img / 1e12 / tf_burst if flop_img else None,
            "loss_first_last": [loss0, loss1], "activation_bytes": plan.act_bytes,
        }
        if args.detail:
            with open(args.detail, "w") as f:
                json.dump({"per_launch_ms": detail, "serial_ms": serial_ms, "kinds": kinds}, f, indent=1)
    # ---- the forward half of BASELINE's metric, in the same line: forward + decode at batch 64 and 256 per GPU ----
    del plan, pool, trainer, host, eng, model
    torch.cuda.empty_cache()
    infer = []
    if not args.no_infer:
        for b in (int(v) for v in args.infer_batches.split(",")):
            r, _ = measure_infer(args, b, dev, rank, world, max(10, args.steps), 3)
            infer.append(r)
            _mark("inference batch %d done (%.0f images/s)" % (b, r["value"]))
    if rank == 0:
        line["infer"] = infer
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_record(args, True)
        print(json.dumps(line))
    _dist_teardown(world)


def run_b200_mhp(args):
    """BASELINE configs[3]: MHP multi-view (4 views per sample) HRNet-W32 + soft-argmax + algebraic DLT triangulation,
    samples sharded over the ranks (all views of a sample on one rank, no collective)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world, dev = _dist_setup()
    from hrnet_b200 import synthetic as fixtures
    from hrnet_b200.models.triangulation import AlgebraicTriangulationNet
    from hrnet_b200.parallel import max_over_ranks
    H, W, B, V = args.height, args.img_width, args.batch, 4
    net = AlgebraicTriangulationNet.from_widths(args.width, image_size=(H, W), device=dev).eval()
    pool = []
    for i in range(3):
        imgs = fixtures.images(B * V, H, W, seed=1 + rank * 16 + i).view(B, V, 3, H, W)
        pool.append((imgs.to(dev), fixtures.cameras(B, V, seed=3 + rank * 16 + i).to(dev)))

    def step(i):
        imgs, P = pool[i % 3]
        with torch.no_grad():
            return net(imgs, P)

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from hrnet_b200 import _lib
    l0 = _lib.lib().hrnb_launch_count()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), device=dev)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)
    # e2e: pinned host images + projection matrices in, 3-D joints out, every step
    host = [(a.cpu().pin_memory(), b.cpu().pin_memory()) for a, b in pool]
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    g0.record()
    for i in range(args.steps):
        imgs, P = host[i % 3]
        with torch.no_grad():
            out = net(imgs.to(dev, non_blocking=True), P.to(dev, non_blocking=True))
        joints = out[0].cpu()
    g1.record()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(g0.elapsed_time(g1), device=dev)
    if rank == 0:
        hbm, tf_burst, tf_sust, peak_kind = peaks()
        fimg = flop_per_img(args.width, H, W)
        line = {"metric": metric_name(args), "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B), "clocks": clocks,
                "e2e": {"value": world * B * args.steps / (e2e_ms / 1e3), "unit": "samples/s",
                        "h2d_bytes_per_step": B * V * 3 * H * W * 4 + B * V * 48, "d2h_bytes_per_step": B * 21 * 3 * 4,
                        "ms_per_step": e2e_ms / args.steps, "api": "AlgebraicTriangulationNet(images [B,4,3,H,W], proj [B,4,3,4]) -> keypoints_3d.cpu()"},
                "gpu_launches": None, "roofline": None, "cpu_baseline": None,
                "images_per_s": value * V,
                "tensor_frac_of_burst_peak": (value * V / world) * fimg / 1e12 / tf_burst if fimg else None}
        print(json.dumps(line))
    _dist_teardown(world)


CONFIGS = {
    1: dict(mode="infer", variant="raw", width=32, height=256, img_width=256, batch=8),
    2: dict(mode="train", variant="softmax", width=32, height=256, img_width=256, batch=64, loss="hm+pose2d"),
    3: dict(mode="infer", variant="softmax", width=48, height=384, img_width=288, batch=256, trainable_temp=True,
            sweep="1,2,4,8,16,32,64,128,256"),       # BASELINE: "batch sweep 1-256"; the headline value is the last (largest) batch
    4: dict(mode="mhp", variant="softmax", width=32, height=256, img_width=256, batch=16, trainable_temp=True),
    5: dict(mode="train", variant="raw", width=48, height=256, img_width=256, batch=32, loss="hm"),
}


def main():
    # NCCL prints its INFO lines to stdout, which must carry exactly one JSON line.  Respect what the launcher asked for
    # (NCCL_DEBUG / NCCL_DEBUG_FILE set by the driver -> its rank check can read the log); if INFO is requested without a
    # file, send it to stderr; only when nothing is set keep NCCL quiet.
    if "NCCL_DEBUG" not in os.environ and "NCCL_DEBUG_FILE" not in os.environ:
        os.environ["NCCL_DEBUG"] = "WARN"
    elif "NCCL_DEBUG_FILE" not in os.environ:
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4, 5], help="BASELINE.json configs[i-1] preset")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--img-width", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=None, choices=["train", "infer", "mhp"])
    ap.add_argument("--variant", default=None, choices=["softmax", "raw"])
    ap.add_argument("--loss", default=None, choices=["hm+pose2d", "hm"])
    ap.add_argument("--trainable-temp", action="store_true", default=None)
    ap.add_argument("--sweep", default=None, help="--mode infer: comma-separated batch sizes, one record each")
    ap.add_argument("--infer-batches", default="64,256", help="--mode train: batch sizes of the forward numbers in `infer`")
    ap.add_argument("--no-infer", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--detail", default=None, help="write per-launch timings (json) here")
    args = ap.parse_args()
    preset = CONFIGS[args.config or 2]
    user_batch = args.batch is not None
    for k, v in dict(dict(loss="hm+pose2d", trainable_temp=False), **preset).items():
        if getattr(args, k, None) is None:
            setattr(args, k, v)
    if user_batch and "sweep" in preset and "--sweep" not in sys.argv:
        args.sweep = None                             # an explicit --batch asks for that one batch, not for the preset's sweep
    if args.variant == "raw":
        args.loss = "hm"
    if args.impl == "reference":
        if args.mode == "mhp":
            print(json.dumps({"impl": "reference", "unavailable": "the multi-view reference arm is not wired (backbone arm: --config 2)"}))
            return
        run_reference(args)
        return
    try:
        {"train": run_b200_train, "infer": run_b200_infer, "mhp": run_b200_mhp}[args.mode](args)
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
