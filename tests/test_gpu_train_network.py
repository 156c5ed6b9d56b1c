"""-m gpu: one whole training step of the CUDA path (forward with batch-stat BN, fused losses, full backward, fused
Adam) against the reference-generated golden vectors (tests/golden/train_*.npz) and the CPU train oracle.

Gradient tolerances: the CUDA path keeps activations and activation-gradients in bf16 (fp32 accumulation, fp32
parameter gradients); the reference is fp32.  Per tensor we bound the relative L2 error and require the cosine with
the reference gradient to be close to 1 (measured values are written to gpurun_out/parity_report.jsonl)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# Whole-network gradients at random init are chaotic in the forward pass: rounding activations to bf16 moves the
# logits of the 2-image golden case by ~35 % (max-rel) and decorrelates early-layer gradients from the fp32 reference
# (cosine 0.3) - measured identically with the CPU oracle when its conv outputs are rounded to bf16, i.e. a property
# of the problem, not of the kernels.  The tests therefore split parity in two:
#   forward  : losses, batch statistics and gradient NORMS against the fp32 reference (golden) within documented bounds;
#   backward : every parameter gradient against the LINEARISED oracle (train_oracle.train_step(conv_values=...)), which
#              differentiates around the activations the CUDA path actually produced - backward is linear given the
#              forward values, so this comparison is tight.
# Measured on B200 (profiles/r1_parity_report.jsonl): worst tensor 0.164 rel-L2 / cosine 0.9865 (stage-2 BN weights, i.e.
# after ~250 bf16-rounded backward ops), head tensors ~1e-2.
LIN_L2_TOL = 0.25       # relative L2 error per parameter tensor vs the linearised oracle (bf16 gradient storage)
LIN_COS_TOL = 0.97


def _report(name, **kv):
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
        f.write(json.dumps(dict(case=name, **kv)) + "\n")


def _setup(variant, trainable, B, H, W, width=32):
    from oracle import fixtures
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet, pose_hrnet_softmax
    cfg = make_cfg(width, softmax=(variant == "softmax"), trainable_softmax=trainable, image_size=(H, W))
    torch.manual_seed(0)
    m = (pose_hrnet_softmax if variant == "softmax" else pose_hrnet).get_pose_net(cfg, is_train=False)
    sd = m.state_dict()
    fixtures.perturb_state_dict(sd)
    m.load_state_dict(sd)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = fixtures.images(B, H, W)
    gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4)
    return m.cuda().train(), cfg, sd, x, gt, xy, vis


def _cmp(got, ref):
    got, ref = got.double().reshape(-1), ref.double().reshape(-1)
    l2 = float((got - ref).norm() / (ref.norm() + 1e-30))
    cos = float((got * ref).sum() / (got.norm() * ref.norm() + 1e-30))
    return l2, cos


@pytest.mark.parametrize("name,variant", [("train_w32_softmax", "softmax"), ("train_w32_raw", "raw")])
def test_train_step_against_reference_golden(golden_dir, name, variant):
    from oracle import fixtures
    from hrnet_b200.train import TrainEngine
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, H, W = int(g["B"]), int(g["H"]), int(g["W"])
    m, cfg, sd, x, gt, xy, vis = _setup(variant, bool(g["trainable_temp"]), B, H, W)
    eng = TrainEngine(m, lr=1e-3, weight_decay=1e-4, loss_factors=(1.0, 0.1), use_graph=False)
    p = eng.train_step(x.cuda(), gt.cuda(), xy.cuda(), vis.cuda(), optimizer_step=False)
    torch.cuda.synchronize()
    losses = p.losses.cpu().numpy()
    assert np.allclose(losses[: (3 if variant == "softmax" else 2)], g["losses"][: (3 if variant == "softmax" else 2)], rtol=2e-2), (losses, g["losses"])
    names = [n for n, _ in m.named_parameters()]
    nat = dict(zip(names, eng.flat.natural_grads()))
    gmax = max(float(np.abs(g["grad/" + str(k)]).max()) for k in g["keys"])
    bad = []
    for k in g["keys"]:
        k = str(k)
        ref = torch.from_numpy(g["grad/" + k])
        if float(ref.abs().max()) <= 1e-4 * gmax:
            continue                                   # mathematically-zero gradients (noise in the reference)
        got = fixtures.sample(nat[k].cpu())
        l2, cos = _cmp(got, ref)
        ratio = float(nat[k].double().norm()) / float(g["gnorm/" + k])
        _report(name + ":" + k, grad_rel_l2=l2, grad_cos=cos, norm_ratio=ratio)
        # chaotic case (2 images, random init, train-mode BN: DESIGN.md §2): gradient NORMS within a factor 2.5 everywhere, head
        # directions pinned; the per-tensor direction checks live in the well-conditioned goldens below
        if not (0.4 < ratio < 2.5) or (k.startswith("last_layer") and cos < 0.85):
            bad.append((k, round(l2, 4), round(cos, 4), round(ratio, 3)))
    assert not bad, bad
    _report(name, loss_total=float(losses[0]), loss_ref=float(g["losses"][0]))
    # running statistics after the step (momentum 0.1, unbiased variance)
    bufs = dict(m.named_buffers())
    for k in ("bn1", "stage3.0.branches.1.2.bn1", "last_layer.1"):
        for s in (".running_mean", ".running_var"):
            ref = g["after/" + k + s]
            got = bufs[k + s].cpu().numpy()
            tol = 2e-2 if k == "bn1" else 0.2          # forward noise grows with depth (see the header comment)
            assert np.abs(got - ref).max() <= tol * max(1e-3, np.abs(ref).max()), (k + s, np.abs(got - ref).max())
    assert int(bufs["bn1.num_batches_tracked"]) == 1
    # optimizer: the fused Adam step on the real parameter layout == torch.optim.Adam fed with the same gradients
    keys = ("conv1.weight", "stage2.0.branches.0.0.conv1.weight", "stage4.2.fuse_layers.3.0.0.0.weight", "last_layer.1.weight",
            "last_layer.3.weight", "last_layer.3.bias")
    before = {k: dict(m.named_parameters())[k].detach().clone() for k in keys}
    grads = {k: nat[k].clone() for k in keys}
    eng.flat.adam_step()
    torch.cuda.synchronize()
    after = dict(m.named_parameters())
    for k in keys:
        ref = torch.nn.Parameter(before[k].clone())
        ref.grad = grads[k]
        torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-4).step()
        assert torch.allclose(after[k].detach(), ref.detach(), rtol=1e-5, atol=1e-7), k


@pytest.mark.parametrize("variant,B,H,W,width", [("softmax", 2, 256, 256, 32), ("raw", 3, 128, 96, 32), ("raw", 2, 128, 128, 48)])
def test_backward_against_linearised_oracle(variant, B, H, W, width):
    """Every parameter gradient of the network against torch autograd differentiating around the CUDA path's own
    forward activations (all conv outputs injected into the CPU oracle)."""
    from oracle import hrnet_oracle, train_oracle
    from hrnet_b200.train import TrainEngine
    m, cfg, sd, x, gt, xy, vis = _setup(variant, True, B, H, W, width=width)
    eng = TrainEngine(m, use_graph=False)
    p = eng.train_step(x.cuda(), gt.cuda(), xy.cuda(), vis.cuda(), optimizer_step=False)
    torch.cuda.synchronize()
    conv_values = {k: c.to_nchw().cpu() for k, c in p.conv_out.items()}
    conv_values["last_layer.3"] = p.out["logits"].cpu()
    o = train_oracle.train_step(sd, x, gt, xy, vis, hrnet_oracle.Arch.from_cfg(cfg), variant, trainable_temp=True,
                                adam=False, conv_values=conv_values)
    assert np.allclose(p.losses.cpu().numpy()[:2], o["losses"][:2], rtol=2e-3), (p.losses, o["losses"])
    names = [n for n, _ in m.named_parameters()]
    nat = dict(zip(names, eng.flat.natural_grads()))
    gmax = max(float(v.abs().max()) for v in o["grads"].values())
    bad, worst = [], (0.0, "", 1.0)
    for k, ref in o["grads"].items():
        if float(ref.abs().max()) <= 1e-4 * gmax:
            continue
        l2, cos = _cmp(nat[k].cpu(), ref)
        if l2 > worst[0]:
            worst = (l2, k, cos)
        if not (l2 < LIN_L2_TOL and cos > LIN_COS_TOL):
            bad.append((k, round(l2, 4), round(cos, 5)))
    _report("train_linearised_w%d_%s_%dx%dx%d" % (width, variant, B, H, W), worst_rel_l2=worst[0], worst_key=worst[1], worst_cos=worst[2],
            n_checked=len(o["grads"]), n_bad=len(bad))
    assert not bad, (len(bad), bad[:20])


def test_graph_replay_equals_eager_and_loss_decreases():
    """CUDA-graph replay of the step gives the same losses as eager launches, and a few Adam steps on a fixed batch
    reduce the loss."""
    from hrnet_b200.train import TrainEngine
    B, H, W = 4, 128, 128
    m, cfg, sd, x, gt, xy, vis = _setup("softmax", False, B, H, W)
    eng = TrainEngine(m, lr=1e-3, use_graph=True)
    xs, gts, xys, viss = x.cuda(), gt.cuda(), xy.cuda(), vis.cuda()
    hist = []
    for it in range(8):
        p = eng.train_step(xs, gts, xys, viss)
        hist.append(float(p.losses[0]))
    assert all(np.isfinite(hist)), hist
    assert hist[-1] < hist[0], hist
    m2, *_ = _setup("softmax", False, B, H, W)
    eng2 = TrainEngine(m2, lr=1e-3, use_graph=False)
    h2 = [float(eng2.train_step(xs, gts, xys, viss).losses[0]) for _ in range(3)]
    assert np.allclose(hist[:3], h2, rtol=5e-3), (hist[:3], h2)


def test_module_train_mode_autograd_matches_fused_step():
    """The drop-in nn.Module in train mode: forward -> reference-style losses (HeatmapLoss + 0.1 * JointsMSELoss on
    get_final_preds) -> loss.backward() -> torch.optim.Adam gives the same gradients / update as the fused train_step."""
    from hrnet_b200.core.loss import HeatmapLoss, JointsMSELoss
    from hrnet_b200.utils.heatmap_decoding import get_final_preds
    from hrnet_b200.train import TrainEngine
    B, H, W = 2, 128, 128
    m, cfg, sd, x, gt, xy, vis = _setup("softmax", True, B, H, W)
    xs, gts, xys, viss = x.cuda(), gt.cuda(), xy.cuda(), vis.cuda()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=1e-3, weight_decay=1e-4)
    heat, feat, temp = m(xs)
    assert heat.shape == (B, 21, H // 4, W // 4) and feat.shape == (B, 480, H // 4, W // 4) and temp is m.trainable_temp
    loss = 1.0 * HeatmapLoss()(heat, gts) + 0.1 * JointsMSELoss()(get_final_preds(heat, True), xys, viss)
    opt.zero_grad()
    loss.backward()
    g_auto = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    assert len(g_auto) == len(list(m.parameters()))
    # fused path on an identical second model
    m2, *_ = _setup("softmax", True, B, H, W)
    eng2 = TrainEngine(m2, use_graph=False)
    p2 = eng2.train_step(xs, gts, xys, viss, optimizer_step=False)
    nat = dict(zip([n for n, _ in m2.named_parameters()], eng2.flat.natural_grads()))
    assert abs(float(loss) - float(p2.losses[0])) <= 1e-4 * abs(float(loss))
    # same kernels on both paths; the heat-map gradient reaches the softmax backward through different (fp32) op orders,
    # and the deepest layers' gradients are cancellation-dominated (cf. the linearised-oracle bounds above)
    gmax = max(float(v.abs().max()) for v in nat.values())
    for n, g in g_auto.items():
        if float(nat[n].abs().max()) <= 1e-4 * gmax:
            continue
        l2, cos = _cmp(g, nat[n])
        assert l2 < 0.15 and cos > 0.99, (n, l2, cos)
    # optimizer step through torch, then a second forward must see the updated weights (re-pack on version change)
    h0 = heat.detach().clone()
    opt.step()
    h1 = m(xs)[0].detach().clone()
    assert float((h1 - h0).abs().max()) > 0                       # the step changed the network
    m.train_engine().repack()                                     # an explicit re-pack must be a no-op now
    h1b = m.train_engine().forward(xs).out["heatmap"]
    torch.cuda.synchronize()
    assert torch.equal(h1, h1b)
    assert int(dict(m.named_buffers())["bn1.num_batches_tracked"]) == 3
    # eval after training uses the updated running statistics (folded inference engine is rebuilt)
    m.eval()
    with torch.no_grad():
        he = m(xs)[0]
    assert torch.isfinite(he).all() and abs(float(he.sum()) - B * 21) < 1e-2 * B * 21


def test_batched_bn_plan_equals_per_unit_plan():
    """horizontally batched BatchNorm launches (default single-stream plan) reproduce the per-unit launches bit for bit:
    same per-tensor block counts, hence the same ordered sums"""
    from hrnet_b200.train import TrainEngine
    B, H, W = 2, 128, 128
    res = []
    for batch in (False, True):
        m, cfg, sd, x, gt, xy, vis = _setup("softmax", True, B, H, W)
        eng = TrainEngine(m, use_graph=batch, bn_batch=batch, multi_stream=False)
        for _ in range(2):
            p = eng.train_step(x.cuda(), gt.cuda(), xy.cuda(), vis.cuda(), optimizer_step=False)
        torch.cuda.synchronize()
        names = [n for n, _ in m.named_parameters()]
        res.append((p.out["logits"].clone(), p.losses.clone(), dict(zip(names, [g.clone() for g in eng.flat.natural_grads()])),
                    {k: v.clone() for k, v in m.named_buffers()}))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    for k in ("bn1.weight", "stage3.1.branches.2.3.bn2.weight", "stage4.2.branches.3.1.bn1.bias", "last_layer.1.bias"):
        assert torch.equal(res[0][2][k], res[1][2][k]), k
    for k in ("stage4.0.branches.1.2.bn2.running_mean", "stage2.0.branches.0.0.bn1.running_var"):
        assert torch.equal(res[0][3][k], res[1][3][k]), k
    a, b = res[0][2]["stage2.0.branches.0.0.conv1.weight"], res[1][2]["stage2.0.branches.0.0.conv1.weight"]
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(a.abs().max()))


def test_multi_stream_plan_equals_single_stream():
    """Branches on separate CUDA streams (hazard-ordered gradient accumulation) must reproduce the single-stream plan:
    forward bit-exactly (ordered reductions), parameter gradients up to the fp32 reduction order of the split-K wgrad."""
    from hrnet_b200.train import TrainEngine
    B, H, W = 3, 128, 128
    res = []
    for multi in (False, True):
        m, cfg, sd, x, gt, xy, vis = _setup("softmax", True, B, H, W)
        eng = TrainEngine(m, use_graph=multi, multi_stream=multi)
        for _ in range(2):                    # second call replays the captured graph in the multi-stream case
            p = eng.train_step(x.cuda(), gt.cuda(), xy.cuda(), vis.cuda(), optimizer_step=False)
        torch.cuda.synchronize()
        res.append((p.out["logits"].clone(), p.losses.clone(), [g.clone() for g in eng.flat.natural_grads()]))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    for a, b in zip(res[0][2], res[1][2]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(a.abs().max()) + 1e-12)


# ---------------------------------------------------------------------------------------------------------------------
# WELL-CONDITIONED training cases: compared with the fp32 reference DIRECTLY (no linearisation).
# The golden files come from the unmodified reference (oracle/make_golden.py train_conditioned_fixture):
#   *_contractive : residual-closing / fuse BatchNorm gammas x 0.05 (fixtures.contract_state_dict) - the random-init net no
#                   longer amplifies bf16 rounding chaotically;
#   *_warm        : the weights after 60 Adam steps of the reference (the trajectory's losses are golden too).
# Bounds are what the bf16 activation / gradient storage delivers on B200 (measured values -> gpurun_out/parity_report.jsonl).
# ---------------------------------------------------------------------------------------------------------------------
# Measured on B200 (profiles/r2_parity_report.jsonl), CUDA path (bf16 activations and activation gradients, fp32 accumulation and
# parameter gradients) against the fp32 reference at identical weights: losses 3e-6; head tensors cosine 0.9995 / rel-L2 0.03;
# the error grows with the number of bf16-stored backward ops a gradient has passed (every BatchNorm backward subtracts the
# mean and the x-hat projection of g, which shrinks the signal but not the rounding noise): stage 4 0.991, stage 2 and stem
# 0.976; over all 920 tensors: cosine min 0.948, 1st percentile 0.960, median 0.989, rel-L2 median 0.149, max 0.318.
COND_COS_MIN = 0.93         # every parameter tensor with a non-negligible gradient
COND_COS_MEDIAN = 0.985     # median over tensors
COND_L2_MEDIAN = 0.20


def _conditioned_model(g, variant):
    from oracle import fixtures
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet, pose_hrnet_softmax
    B, H, W, width = int(g["B"]), int(g["H"]), int(g["W"]), int(g["width"])
    cfg = make_cfg(width, softmax=(variant == "softmax"), trainable_softmax=True, image_size=(H, W))
    torch.manual_seed(0)
    m = (pose_hrnet_softmax if variant == "softmax" else pose_hrnet).get_pose_net(cfg, is_train=False)
    sd = m.state_dict()
    fixtures.perturb_state_dict(sd)
    if float(g["contract"]) > 0:
        fixtures.contract_state_dict(sd, float(g["contract"]))
    m.load_state_dict(sd)
    return m, cfg, {k: v.clone() for k, v in m.state_dict().items()}, (B, H, W)


def _warm_oracle(g, cfg, sd, variant, B, H, W):
    """the reference's warm-up trajectory re-run with the (pinned) oracle port on the host; checked against the golden
    losses and weight checksums before anything is compared with it"""
    from oracle import hrnet_oracle, train_oracle
    from oracle.make_golden import warm_batch
    arch = hrnet_oracle.Arch.from_cfg(cfg)
    ostate, losses = None, []
    for s in range(int(g["warm_steps"])):
        x, gt, xy, vis = warm_batch(s, B, H, W)
        o = train_oracle.train_step(sd, x, gt, xy, vis, arch, variant, trainable_temp=True, opt_state=ostate)
        sd, ostate = o["state"], o["opt_state"]
        losses.append(o["losses"])
    if losses:
        # two fp32 runs of this trajectory do not stay together (the UNMODIFIED reference with 3 instead of 8 host threads is
        # 1.5 % off in the pose2d loss after 16 steps, profiles/r2_reference_self_chaos.txt): loose check on the port's run
        rel = np.abs(np.array(losses) - g["trajectory"]) / np.abs(g["trajectory"])
        assert rel[:, 0].max() < 2e-2 and rel[:, 2].max() < 0.15, ("oracle warm-up trajectory left the reference's", rel.max(0))
        assert np.allclose(losses[0], g["trajectory"][0], rtol=1e-4)      # step 0 (identical weights) is tight
    return sd


@pytest.mark.parametrize("name,variant", [("train_w32_softmax_contractive", "softmax"), ("train_w32_softmax_warm", "softmax"),
                                          ("train_w48_raw_contractive", "raw")])
def test_conditioned_gradients_against_fp32_reference(golden_dir, name, variant):
    from oracle import fixtures, hrnet_oracle, train_oracle
    from hrnet_b200.train import TrainEngine
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m, cfg, sd, (B, H, W) = _conditioned_model(g, variant)
    sd = _warm_oracle(g, cfg, sd, variant, B, H, W)
    m.load_state_dict(sd)
    warm = int(g["warm_steps"]) > 0
    if not warm:
        for k in g["keys"]:                  # the weights under test are exactly the reference's (sampled check)
            k = str(k)
            assert np.allclose(fixtures.sample(sd[k]).numpy(), g["weight/" + k], rtol=1e-6, atol=1e-8), k
    x = fixtures.images(B, H, W)
    gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4)
    eng = TrainEngine(m.cuda().train(), use_graph=False)
    p = eng.train_step(x.cuda(), gt.cuda(), xy.cuda(), vis.cuda(), optimizer_step=False)
    torch.cuda.synchronize()
    losses = p.losses.cpu().numpy()
    nl = 3 if variant == "softmax" else 2
    if not warm:
        assert np.allclose(losses[:nl], g["losses"][:nl], rtol=5e-3), (losses, g["losses"])
    names = [n for n, _ in m.named_parameters()]
    nat = dict(zip(names, eng.flat.natural_grads()))
    # (1) the sampled gradients of the UNMODIFIED reference (identical weights: the cases without a warm-up)
    gmax = float(g["gmax_all"].max())
    stats = []
    for k in ([] if warm else g["keys"]):
        k = str(k)
        ref = torch.from_numpy(g["grad/" + k])
        if float(ref.abs().max()) <= 1e-4 * gmax:
            continue
        l2, cos = _cmp(fixtures.sample(nat[k].cpu()), ref)
        stats.append((cos, l2, k))
        _report(name + ":" + k, grad_rel_l2=l2, grad_cos=cos)
    # (2) every parameter tensor against the fp32 oracle step at the same weights (the oracle is pinned to the reference at
    #     exactly these weights by make_golden: worst cosine > 0.999) - NOT linearised: plain fp32 forward and backward
    o = train_oracle.train_step(sd, x, gt, xy, vis, hrnet_oracle.Arch.from_cfg(cfg), variant, trainable_temp=True, adam=False)
    norms = dict(zip([str(k) for k in g["all_keys"]], g["gnorm_all"]))
    allstats = []
    for k, ref in o["grads"].items():
        if float(ref.abs().max()) <= 1e-4 * gmax:
            continue                      # mathematically-zero gradients (bias before BatchNorm, softmax shift invariance)
        if not warm:
            assert abs(float(ref.double().norm()) - norms[k]) <= 2e-3 * norms[k] + 1e-12, ("oracle != reference gradient norm", k)
        l2, cos = _cmp(nat[k].cpu(), ref)
        allstats.append((cos, l2, k))
    allstats.sort()
    cs = np.array([c for c, _, _ in allstats])
    ls = np.array([l for _, l, _ in allstats])
    _report(name, n_tensors=len(allstats), cos_min=float(cs.min()), cos_p01=float(np.percentile(cs, 1)), cos_median=float(np.median(cs)),
            frac_cos_ge_099=float((cs >= 0.99).mean()), l2_median=float(np.median(ls)), l2_max=float(ls.max()),
            worst=[(round(c, 4), round(l, 3), k) for c, l, k in allstats[:5]], loss=losses.tolist(), loss_ref=g["losses"].tolist())
    # warm case (measured on B200): median cosine 0.9937 / rel-L2 0.114, 60 % of the tensors >= 0.99, but a longer tail
    # (min 0.888, 1st percentile 0.918: BatchNorm biases of the 64x64 branch whose gradients are sums of 32k nearly cancelling terms)
    cos_min = 0.85 if warm else COND_COS_MIN
    assert cs.min() >= cos_min, allstats[:8]
    assert np.median(cs) >= COND_COS_MEDIAN and np.median(ls) <= COND_L2_MEDIAN, (float(np.median(cs)), float(np.median(ls)))
    assert np.allclose(losses[:nl], np.array(o["losses"])[:nl], rtol=5e-3), (losses, o["losses"])
    if stats:
        assert min(c for c, _, _ in stats) >= COND_COS_MIN, sorted(stats)[:5]


def test_loss_trajectory_against_reference(golden_dir):
    """60 fused training steps (forward, losses, backward, Adam, re-pack; CUDA graph replays) from the reference's initial
    weights on the reference's batches: the three loss curves must follow the UNMODIFIED reference's (golden trajectory of
    train_w32_softmax_warm.npz) and the weights must end up where the reference's did."""
    from oracle import fixtures
    from oracle.make_golden import warm_batch
    from hrnet_b200.train import TrainEngine
    g = np.load(os.path.join(golden_dir, "train_w32_softmax_warm.npz"))
    m, cfg, sd, (B, H, W) = _conditioned_model(g, "softmax")
    eng = TrainEngine(m.cuda().train(), lr=1e-3, weight_decay=1e-4, loss_factors=(1.0, 0.1), use_graph=True)
    traj = []
    for s in range(int(g["warm_steps"])):
        x, gt, xy, vis = warm_batch(s, B, H, W)
        p = eng.train_step(x.cuda(), gt.cuda(), xy.cuda(), vis.cuda())
        traj.append(p.losses.cpu().numpy().copy())
    traj = np.array(traj, dtype=np.float64)
    ref = g["trajectory"]
    rel = np.abs(traj - ref) / np.abs(ref)
    _report("trajectory_w32_softmax_60_steps", max_rel_total=float(rel[:, 0].max()), max_rel_hm=float(rel[:, 1].max()),
            max_rel_p2d=float(rel[:, 2].max()), first=traj[0].tolist(), last=traj[-1].tolist(), ref_last=ref[-1].tolist())
    assert rel[:, 0].max() < 2e-2 and rel[:, 1].max() < 2e-2, (rel[:, 0].max(), rel[:, 1].max())
    # the pose2d term alone (10 % of the total): two fp32 runs of the UNMODIFIED reference (8 vs 3 host threads) already
    # differ by 1.5 % after 16 steps (profiles/r2_reference_self_chaos.txt)
    assert rel[:, 2].max() < 0.15, rel[:, 2].max()
    # the weights after 60 steps: Adam's first steps move every element by ~lr whatever the gradient's size, so the UPDATE
    # direction (w_60 - w_0) is compared, and against the floor two fp32 runs set: the golden stores the cosine between the
    # reference's update and the oracle port's (a second fp32 run of the same 60 steps): 0.55 (stem) ... 0.71 (last_layer.3) ...
    # 0.98 (stage 4).  Measured for the bf16 path on B200: 0.48 (last_layer.3) / 0.44 (last_layer.0) / 0.67 (stage-4 fuse conv) /
    # 0.30 (8x8-branch conv, floor 0.98: Adam's per-element normalisation gives the many small, bf16-noisy gradient elements of
    # that tensor the same weight as the few large ones - its magnitude-weighted gradient cosine is 0.986 in the contractive
    # case; 0.15-0.22 for a stage-2 conv, floor 0.74; 0.07 for the stem's conv2, floor 0.56 - the deeper the backward chain, the
    # sooner the bf16 rounding noise (1e-3 against the 1e-6 that separates the two fp32 runs) saturates the chaotic divergence).
    # All are REPORTED (gpurun_out/parity_report.jsonl -> profiles/); asserted are the head / stage-4 tensors, whose updates stay
    # clearly correlated with the reference's (> 0.25) - the early layers are covered by the direct gradient comparisons of the
    # contractive and warm cases above, not by this 60-step chaotic roll-out.
    cur = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    for k in ("last_layer.3.weight", "last_layer.0.weight", "stage4.2.fuse_layers.0.3.0.weight", "stage4.0.branches.3.0.conv1.weight",
              "stage2.0.branches.0.0.conv1.weight", "conv2.weight"):
        d_ref = torch.from_numpy(g["weight/" + k]) - fixtures.sample(sd[k])
        d_got = fixtures.sample(cur[k]) - fixtures.sample(sd[k])
        l2, cos = _cmp(d_got, d_ref)
        floor = float(g["update_cos_floor/" + k])
        _report("trajectory_update:" + k, rel_l2=l2, cos=cos, fp32_floor=floor)
        if k.startswith("last_layer") or k.startswith("stage4"):
            assert cos > (0.25 if k.startswith("last_layer") else 0.15), (k, l2, cos, floor)   # measured 0.43-0.52 / 0.30-0.72
