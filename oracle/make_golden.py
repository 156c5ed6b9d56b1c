"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the authoring container only:   python oracle/make_golden.py
The fixtures pin (a) the oracle restatements and (b) the product's seeded parameter initialisation
against the reference itself; tests/test_oracle_golden.py re-checks them without needing the reference.
TEST INFRASTRUCTURE.
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, fixtures, hrnet_oracle, decode_oracle, loss_oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CHECK_KEYS = ["conv1.weight", "layer1.0.downsample.0.weight", "layer1.3.conv3.weight", "transition1.1.0.0.weight",
              "stage2.0.fuse_layers.1.0.0.0.weight", "stage3.2.branches.2.1.conv2.weight",
              "stage4.2.fuse_layers.3.0.2.0.weight", "stage4.1.fuse_layers.0.3.0.weight", "last_layer.0.weight",
              "last_layer.0.bias", "last_layer.3.weight", "last_layer.3.bias"]


def sub(t, cs=1, ss=2):
    return t[:, ::cs, ::ss, ::ss].contiguous().numpy()


def net_fixture(name, yaml_rel, variant, width_override=None, H=256, W=256, B=1, sharpen=False, perturb=True):
    pose_hrnet, pose_hrnet_softmax, *_ = ref_shim.modules()
    cfg = ref_shim.load_cfg(yaml_rel)
    if width_override:
        for s, nb in ((2, 2), (3, 3), (4, 4)):
            cfg.MODEL.EXTRA["STAGE%d" % s]["NUM_CHANNELS"] = [width_override * 2 ** i for i in range(nb)]
    mod = pose_hrnet_softmax if variant == "softmax" else pose_hrnet
    torch.manual_seed(0)
    ref = mod.get_pose_net(cfg, is_train=False).eval()
    sd0 = ref.state_dict()
    init_sums = fixtures.tensor_checksums(sd0, CHECK_KEYS)
    total_abs = float(sum(v.double().abs().sum() for k, v in sd0.items() if v.dtype.is_floating_point))
    if perturb:
        fixtures.perturb_state_dict(sd0)
    if sharpen:
        fixtures.sharpen_head(sd0)
    if variant == "softmax" and perturb:
        sd0["trainable_temp"].fill_(1.7)
    ref.load_state_dict(sd0)
    x = fixtures.images(B, H, W)
    torch.set_num_threads(8)
    with torch.no_grad():
        out = ref(x)
    arch = hrnet_oracle.Arch.from_cfg(cfg)
    o = hrnet_oracle.forward(ref.state_dict(), x, arch, variant)
    rec = {"init_keys": np.array(CHECK_KEYS), "init_sums": np.array([init_sums[k] for k in CHECK_KEYS]),
           "init_total_abs": np.array(total_abs), "n_keys": np.array(len(sd0)),
           "key_list_hash": np.array(hash_keys(sd0)), "H": np.array(H), "W": np.array(W), "B": np.array(B)}
    if variant == "softmax":
        heat, feat, temp = out
        assert torch.allclose(o[0], heat, rtol=1e-4, atol=1e-9), (o[0] - heat).abs().max()
        assert torch.allclose(o[1], feat, rtol=1e-4, atol=1e-5)
        rec.update(heat=sub(heat), feat=sub(feat, 16, 4), logits=sub(o[3]), temp=np.array(float(temp)),
                   heat_sum=np.array(float(heat.double().sum())), feat_abs=np.array(float(feat.double().abs().sum())))
        # decode through the reference's own functions
        _, _, hd, inf, _ = ref_shim.modules()
        rec["soft_coords"] = hd.get_final_preds(heat, True).numpy()
        rec["argmax_h"] = hd.get_final_preds(heat, False).numpy()
        p, mv = inf.get_max_preds(heat.numpy())
        rec["max_preds"], rec["maxvals"] = p, mv
    else:
        logits, feat = out
        assert torch.allclose(o[0], logits, rtol=1e-4, atol=1e-6), (o[0] - logits).abs().max()
        assert torch.allclose(o[1], feat, rtol=1e-4, atol=1e-5)
        rec.update(logits=sub(logits), feat=sub(feat, 4, 4),
                   logits_abs=np.array(float(logits.double().abs().sum())),
                   feat_abs=np.array(float(feat.double().abs().sum())))
        _, _, hd, inf, _ = ref_shim.modules()
        p, mv = inf.get_max_preds(logits.numpy())
        rec["max_preds"], rec["maxvals"] = p, mv
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(name, "ok; oracle == reference; keys", len(sd0))


def hash_keys(sd):
    import hashlib
    return hashlib.sha256("\n".join(sorted("%s %s %s" % (k, tuple(v.shape), v.dtype) for k, v in sd.items())).encode()).hexdigest()


def decode_fixture():
    _, _, hd, inf, _ = ref_shim.modules()
    g = torch.Generator().manual_seed(7)
    cases = {}
    for tag, (B, J, h, w) in {"sq": (2, 21, 16, 16), "rect": (2, 21, 24, 16), "j20": (1, 20, 8, 12)}.items():
        logits = torch.randn(B, J, h, w, generator=g) * 3
        hm = torch.softmax(logits.reshape(B, J, -1), 2).reshape(B, J, h, w)
        edge = logits.clone()
        edge[0, 0] = 0.0                                  # all-zero map
        edge[0, 1] = -edge[0, 1].abs() - 1.0              # all-negative map
        edge[0, 2] = 0.0; edge[0, 2, 3, 5] = 2.0; edge[0, 2, 5, 1] = 2.0       # exact tie: first wins
        for k, (py, px) in enumerate([(0, 0), (1, 1), (h - 2, w - 2), (h - 1, w - 1), (2, 2), (h - 3, w - 3), (2, w - 2)]):
            edge[0, 3 + k] = torch.rand(h, w, generator=g) * 0.1
            edge[0, 3 + k, py, px] = 5.0
        edge[0, 10] = 0.0; edge[0, 10, 4, 4] = 1.0        # flat neighbourhood -> sign(0) = 0
        center = (torch.rand(B, 2, generator=g) * 200 + 60).numpy().astype(np.float32)
        scale = (torch.rand(B, 2, generator=g) + 0.5).numpy().astype(np.float32)
        for nm, t in (("soft", hm), ("edge", edge)):
            a = t.numpy()
            cases["%s_%s_in" % (tag, nm)] = a
            p, mv = inf.get_max_preds(a.copy())
            cases["%s_%s_maxpreds" % (tag, nm)], cases["%s_%s_maxvals" % (tag, nm)] = p, mv
            cases["%s_%s_hstride" % (tag, nm)] = hd.get_final_preds(t, False).numpy()
            cases["%s_%s_expect" % (tag, nm)] = hd.get_final_preds(t, True).numpy()
            for pp in (0, 1):
                cfg = ref_shim.to_attr({"TEST": {"POST_PROCESS": bool(pp)}})
                fp, fmv = inf.get_final_preds(cfg, a.copy(), center, scale)
                cases["%s_%s_final%d" % (tag, nm, pp)] = fp
            # oracle checks
            op, omv = decode_oracle.get_max_preds(a.copy())
            assert np.array_equal(op, p) and np.array_equal(omv, mv)
            assert np.array_equal(decode_oracle.argmax_hstride(a), cases["%s_%s_hstride" % (tag, nm)])
            assert np.allclose(decode_oracle.spatial_expectation2d(a), cases["%s_%s_expect" % (tag, nm)], rtol=1e-4, atol=1e-3), np.abs(decode_oracle.spatial_expectation2d(a) - cases["%s_%s_expect" % (tag, nm)]).max()
            for pp in (0, 1):
                ofp, _ = decode_oracle.final_preds(a.copy(), center, scale, bool(pp))
                assert np.allclose(ofp, cases["%s_%s_final%d" % (tag, nm, pp)], rtol=1e-5, atol=1e-4), tag
        cases[tag + "_center"], cases[tag + "_scale"] = center, scale
        cases[tag + "_logits"] = logits.numpy()
        sm_ref = torch.softmax(logits.reshape(B, J, -1) * 1.7, 2).reshape(B, J, h, w).numpy()
        assert np.allclose(decode_oracle.spatial_softmax(logits.numpy(), 1.7), sm_ref, rtol=1e-5, atol=1e-8)
        cases[tag + "_softmax17"] = sm_ref
    np.savez_compressed(os.path.join(GOLD, "decode.npz"), **cases)
    print("decode ok; oracle == reference")


def loss_fixture():
    *_, loss = ref_shim.modules()
    g = torch.Generator().manual_seed(11)
    rec = {}
    for tag, (B, J, h, w) in {"a": (2, 21, 16, 16), "b": (3, 20, 8, 12)}.items():
        pred = torch.rand(B, J, h, w, generator=g, requires_grad=True)
        gt, xy, vis = fixtures.targets(B, J, h, w)
        for mode in ("l2", "l1"):
            l = loss.HeatmapLoss(mode)(pred, gt)
            (gr,) = torch.autograd.grad(l, pred)
            rec["%s_hm_%s" % (tag, mode)] = l.detach().numpy()
            rec["%s_hm_%s_grad" % (tag, mode)] = gr.numpy()
            assert np.allclose(loss_oracle.heatmap_loss(pred.detach().numpy(), gt.numpy(), mode), l.item(), rtol=1e-5)
            assert np.allclose(loss_oracle.heatmap_loss_grad(pred.detach().numpy(), gt.numpy(), mode), gr.numpy(), rtol=1e-5, atol=1e-8)
        pp = (xy + torch.randn(B, J, 2, generator=g) * 2).requires_grad_(True)
        with torch.no_grad():
            pp[0, 0] = xy[0, 0]                      # exactly-zero distance -> zero gradient
        for vtag, v in (("vis", vis), ("novis", None), ("zerovis", torch.zeros_like(vis))):
            l = loss.JointsMSELoss()(pp, xy, v)
            (gr,) = torch.autograd.grad(l, pp)
            rec["%s_p2d_%s" % (tag, vtag)] = l.detach().numpy()
            rec["%s_p2d_%s_grad" % (tag, vtag)] = gr.numpy()
            vn = v.numpy() if v is not None else None
            assert np.allclose(loss_oracle.pose2d_loss(pp.detach().numpy(), xy.numpy(), vn), l.item(), rtol=1e-5)
            assert np.allclose(loss_oracle.pose2d_loss_grad(pp.detach().numpy(), xy.numpy(), vn), gr.numpy(), rtol=1e-4, atol=1e-7)
        rec.update({tag + "_pred": pred.detach().numpy(), tag + "_gt": gt.numpy(), tag + "_pp": pp.detach().numpy(),
                    tag + "_xy": xy.numpy(), tag + "_vis": vis.numpy()})
    np.savez_compressed(os.path.join(GOLD, "loss.npz"), **rec)
    print("loss ok; oracle == reference")


TRAIN_KEYS = ["conv1.weight", "bn1.weight", "conv2.weight", "layer1.0.downsample.0.weight", "layer1.0.conv2.weight",
              "layer1.3.bn3.bias", "transition1.0.0.weight", "transition1.1.0.0.weight", "stage2.0.branches.0.0.conv1.weight",
              "stage2.0.fuse_layers.0.1.0.weight", "stage2.0.fuse_layers.1.0.0.0.weight", "stage3.1.branches.2.3.conv2.weight",
              "stage3.3.fuse_layers.2.0.1.0.weight", "stage4.0.branches.3.0.conv1.weight", "stage4.2.fuse_layers.0.3.0.weight",
              "stage4.2.fuse_layers.3.0.0.0.weight", "stage4.2.fuse_layers.0.2.1.weight", "last_layer.0.weight",
              "last_layer.1.weight", "last_layer.3.weight", "last_layer.3.bias"]


def train_fixture(name, yaml_rel, variant, B=2, H=256, W=256, trainable_temp=False):
    """One training step of the UNMODIFIED reference (model.train(), its own losses / decode, torch Adam as
    utils.get_optimizer builds it): losses, gradients, running stats and updated parameters."""
    from oracle import train_oracle
    pose_hrnet, pose_hrnet_softmax, hd, _, loss = ref_shim.modules()
    cfg = ref_shim.load_cfg(yaml_rel)
    if variant == "softmax":
        cfg.MODEL["TRAINABLE_SOFTMAX"] = trainable_temp
    mod = pose_hrnet_softmax if variant == "softmax" else pose_hrnet
    torch.manual_seed(0)
    ref = mod.get_pose_net(cfg, is_train=False)
    sd0 = copy.deepcopy(ref.state_dict())
    fixtures.perturb_state_dict(sd0)          # non-trivial gamma / beta
    ref.load_state_dict(sd0)
    ref.train()
    x = fixtures.images(B, H, W)
    gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4)
    torch.set_num_threads(8)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, ref.parameters()), lr=1e-3, weight_decay=1e-4)
    out = ref(x)
    if variant == "softmax":
        heat = out[0]
        l_hm = loss.HeatmapLoss()(heat, gt)
        l_p2d = loss.JointsMSELoss()(hd.get_final_preds(heat, True), xy, vis)
        total = 1.0 * l_hm + 0.1 * l_p2d
    else:
        l_hm = loss.HeatmapLoss()(out[0], gt)
        l_p2d = torch.zeros(())
        total = 1.0 * l_hm
    opt.zero_grad()
    total.backward()
    grads = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    opt.step()
    after = ref.state_dict()
    arch = hrnet_oracle.Arch.from_cfg(cfg)
    o = train_oracle.train_step(sd0, x, gt, xy, vis, arch, variant, trainable_temp=trainable_temp)
    assert np.allclose(o["losses"], (float(total), float(l_hm), float(l_p2d)), rtol=1e-5), (o["losses"], float(total))
    worst = 0.0
    gmax = max(float(g.abs().max()) for g in grads.values())
    for k, g in grads.items():
        og = o["grads"][k]
        # (a conv bias followed by BatchNorm has a mathematically zero gradient: only rounding noise, hence the floor)
        err = float((og - g).abs().max() / max(float(g.abs().max()), 1e-4 * gmax))
        worst = max(worst, err)
        assert err < 2e-3, (k, err)
    for k, v in after.items():
        if v.dtype.is_floating_point:
            # the first Adam step moves every element by ~lr*sign(g): elements whose gradient is rounding noise may
            # flip, so the check is "all within 2*lr + almost all identical"
            d = (o["state"][k] - v).abs()
            assert float(d.max()) <= 2.1e-3 and float((d > 1e-5).float().mean()) < 0.02, (k, float(d.max()), float((d > 1e-5).float().mean()))
    rec = {"losses": np.array([float(total), float(l_hm), float(l_p2d)]), "B": np.array(B), "H": np.array(H), "W": np.array(W),
           "keys": np.array(TRAIN_KEYS), "trainable_temp": np.array(int(trainable_temp))}
    for k in TRAIN_KEYS:          # big tensors are stored as a strided sample (fixtures.sample) + their L2 norm
        rec["grad/" + k] = fixtures.sample(grads[k]).numpy()
        rec["gnorm/" + k] = np.array(float(grads[k].double().norm()))
        rec["after/" + k] = fixtures.sample(after[k]).numpy()
    for k in ("bn1", "stage3.0.branches.1.2.bn1", "last_layer.1"):
        rec["after/" + k + ".running_mean"] = after[k + ".running_mean"].numpy()
        rec["after/" + k + ".running_var"] = after[k + ".running_var"].numpy()
    rec["grad_l2_all"] = np.array([float(g.double().norm()) for _, g in sorted(grads.items())])
    if "trainable_temp" in grads:
        rec["grad/trainable_temp"] = grads["trainable_temp"].numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(name, "ok; oracle train step == reference; worst grad rel err %.2e; losses" % worst, rec["losses"])


WARM_STEPS = 60


def warm_batch(step, B, H, W):
    """batch `step` of the warm-up trajectory (shared by make_golden.py and tests/test_gpu_train_network.py)"""
    x = fixtures.images(B, H, W, seed=1000 + step)
    gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4, seed=2000 + step)
    return x, gt, xy, vis


def _ref_losses(ref, loss, hd, variant, x, gt, xy, vis):
    out = ref(x)
    if variant == "softmax":
        heat = out[0]
        l_hm = loss.HeatmapLoss()(heat, gt)
        l_p2d = loss.JointsMSELoss()(hd.get_final_preds(heat, True), xy, vis)
        return 1.0 * l_hm + 0.1 * l_p2d, l_hm, l_p2d
    l_hm = loss.HeatmapLoss()(out[0], gt)
    return 1.0 * l_hm, l_hm, torch.zeros(())


def train_conditioned_fixture(name, yaml_rel, variant, B=8, H=128, W=128, warm_steps=0, contract=None, width_override=None):
    """A WELL-CONDITIONED training case, straight from the UNMODIFIED reference: (optionally) `warm_steps` Adam steps
    of the reference on seeded batches (their losses are the golden TRAJECTORY), then one more forward/backward on a
    fixed batch whose gradients are the golden gradients.  With `contract` the residual / fuse BatchNorm gammas are
    scaled down instead (fixtures.contract_state_dict).  The oracle port is run alongside and asserted equal."""
    from oracle import train_oracle
    pose_hrnet, pose_hrnet_softmax, hd, _, loss = ref_shim.modules()
    cfg = ref_shim.load_cfg(yaml_rel)
    if width_override:
        for s_, nb in ((2, 2), (3, 3), (4, 4)):
            cfg.MODEL.EXTRA["STAGE%d" % s_]["NUM_CHANNELS"] = [width_override * 2 ** i for i in range(nb)]
    if variant == "softmax":
        cfg.MODEL["TRAINABLE_SOFTMAX"] = True
    mod = pose_hrnet_softmax if variant == "softmax" else pose_hrnet
    torch.manual_seed(0)
    ref = mod.get_pose_net(cfg, is_train=False)
    sd0 = copy.deepcopy(ref.state_dict())
    fixtures.perturb_state_dict(sd0)
    if contract:
        fixtures.contract_state_dict(sd0, contract)
    ref.load_state_dict(sd0)
    ref.train()
    arch = hrnet_oracle.Arch.from_cfg(cfg)
    torch.set_num_threads(8)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, ref.parameters()), lr=1e-3, weight_decay=1e-4)
    traj, max_dev = [], 0.0
    osd = {k: v.clone() for k, v in sd0.items()}
    ostate = None
    for s_ in range(warm_steps):
        x, gt, xy, vis = warm_batch(s_, B, H, W)
        total, l_hm, l_p2d = _ref_losses(ref, loss, hd, variant, x, gt, xy, vis)
        opt.zero_grad()
        total.backward()
        opt.step()
        traj.append([float(total), float(l_hm), float(l_p2d)])
        o = train_oracle.train_step(osd, x, gt, xy, vis, arch, variant, trainable_temp=True, opt_state=ostate)
        osd, ostate = o["state"], o["opt_state"]
        dev = float(np.max(np.abs(np.array(o["losses"]) - np.array(traj[-1])) / np.abs(np.array(traj[-1]))))
        max_dev = max(max_dev, dev)
        assert dev < 0.15, (s_, o["losses"], traj[-1])        # chaos floor, see below; a port bug would show at step 0
        if s_ % 10 == 0:
            print("  warm step", s_, traj[-1], "oracle-vs-reference loss deviation so far %.2e" % max_dev, flush=True)
    # the golden step: fixed batch, gradients only (no optimizer step)
    x = fixtures.images(B, H, W)
    gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4)
    total, l_hm, l_p2d = _ref_losses(ref, loss, hd, variant, x, gt, xy, vis)
    opt.zero_grad()
    total.backward()
    grads = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    cur = {k: v.clone() for k, v in ref.state_dict().items()}
    # pin the oracle port AT THE REFERENCE'S (warm) WEIGHTS: same losses, same gradients.  (Its own 60-step trajectory,
    # run alongside above, drifts from the reference's like any second fp32 run does - the unmodified reference with 3
    # instead of 8 threads is 1.5 % off in the pose2d loss after 16 steps, profiles/r2_reference_self_chaos.txt - so
    # trajectories are compared with the tolerances that noise floor allows, single steps tightly.)
    o = train_oracle.train_step(cur, x, gt, xy, vis, arch, variant, trainable_temp=True, adam=False)
    assert np.allclose(o["losses"], (float(total), float(l_hm), float(l_p2d)), rtol=2e-5), (o["losses"], float(total))
    gmax = max(float(g.abs().max()) for g in grads.values())
    worst_cos, worst_err = 1.0, 0.0
    for k, g in grads.items():
        if float(g.abs().max()) <= 1e-4 * gmax:
            continue
        a, b = o["grads"][k].double().reshape(-1), g.double().reshape(-1)
        worst_cos = min(worst_cos, float((a * b).sum() / (a.norm() * b.norm() + 1e-30)))
        worst_err = max(worst_err, float((a - b).abs().max() / b.abs().max()))
    print("  oracle vs reference at the reference's weights: worst gradient cosine %.8f, worst max-rel error %.2e; "
          "trajectory loss deviation of the oracle's own run %.2e" % (worst_cos, worst_err, max_dev), flush=True)
    assert worst_cos > 0.99999 and worst_err < 1e-2, (worst_cos, worst_err)
    rec = {"losses": np.array([float(total), float(l_hm), float(l_p2d)]), "B": np.array(B), "H": np.array(H), "W": np.array(W),
           "keys": np.array(TRAIN_KEYS), "warm_steps": np.array(warm_steps), "contract": np.array(contract or 0.0),
           "trajectory": np.array(traj, dtype=np.float64).reshape(-1, 3), "width": np.array(width_override or 32),
           "oracle_vs_ref_traj_dev": np.array(max_dev), "oracle_vs_ref_worst_cos": np.array(worst_cos)}
    for k in TRAIN_KEYS:
        rec["grad/" + k] = fixtures.sample(grads[k]).numpy()
        rec["weight/" + k] = fixtures.sample(cur[k]).numpy()
        if warm_steps:
            # fp32 floor of "the weights end up where the reference's did": cosine between the 60-step weight UPDATES of the
            # reference and of the oracle port (a second fp32 run of the same trajectory)
            da, db = (osd[k] - sd0[k]).double().reshape(-1), (cur[k] - sd0[k]).double().reshape(-1)
            rec["update_cos_floor/" + k] = np.array(float((da * db).sum() / (da.norm() * db.norm() + 1e-30)))
    names = sorted(grads)
    rec["all_keys"] = np.array(names)
    rec["gnorm_all"] = np.array([float(grads[k].double().norm()) for k in names])
    rec["gmax_all"] = np.array([float(grads[k].abs().max()) for k in names])
    rec["wsum_all"] = np.array([float(cur[k].double().sum()) for k in names])
    if "trainable_temp" in grads:
        rec["grad/trainable_temp"] = grads["trainable_temp"].numpy()
        rec["temp"] = np.array(float(cur["trainable_temp"]))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(name, "ok; oracle == reference over %d warm-up steps; worst oracle-vs-reference gradient cosine %.6f; losses" % (warm_steps, worst_cos), rec["losses"])


def triangulation_fixture():
    """the UNMODIFIED lib/utils/misc.py DLT_sii_pytorch, called per joint exactly like
    lib/models/triangulation.py:258-261, with the removed torch.solve(b, A) mapped onto torch.linalg.solve(A, b)."""
    from oracle import triangulation_oracle as T
    ref_shim.install()
    import utils.misc as M
    torch.solve = lambda b, A: (torch.linalg.solve(A, b), None)     # the only shim: argument order of the removed API
    rec = {}
    for name, (B, V, J, seed, noise) in {"mhp4": (6, 4, 21, 11, 1.5), "two_views": (3, 2, 21, 12, 0.5),
                                         "eight_views_j20": (2, 8, 20, 13, 3.0)}.items():
        P = fixtures.cameras(B, V, seed=seed)
        X, uv = fixtures.multiview_joints(P, J, seed=seed + 100, noise_px=noise)
        torch.manual_seed(seed)
        ref = torch.cat([M.DLT_sii_pytorch(uv[:, :, k].clone(), P.clone()).unsqueeze(1) for k in range(J)], dim=1)
        bk0 = T.start_vectors(B, J, seed)
        ours = T.triangulate_joints(uv.numpy(), P.numpy(), bk0)
        err = np.abs(ours - ref.numpy()).max() / np.abs(ref.numpy()).max()
        assert err < 2e-5, (name, err)
        svd = torch.stack([M.triangulate_from_multiple_views_svd(P.clone(), uv[:, :, k].clone()) for k in range(J)], dim=1)
        assert np.abs(T.svd_triangulation(uv[:, :, 0].numpy(), P.numpy()) - svd[:, 0].numpy()).max() < 1e-2
        # the reference's own autograd through DLT_sii_pytorch (train3D trains the backbone through it): d loss / d points
        gw = torch.randn(B, J, 3, generator=torch.Generator().manual_seed(seed + 500))
        uvg = uv.clone().requires_grad_(True)
        torch.manual_seed(seed)
        refg = torch.cat([M.DLT_sii_pytorch(uvg[:, :, k], P.clone()).unsqueeze(1) for k in range(J)], dim=1)
        (refg * gw).sum().backward()
        d_uv = uvg.grad.numpy()
        mine = np.stack([T.dlt_sii_backward(uv[:, :, k].numpy(), P.numpy(), bk0[k], gw[:, k].numpy()) for k in range(J)], axis=2)
        gerr = np.abs(mine - d_uv).max() / np.abs(d_uv).max()
        assert gerr < 2e-3, (name, gerr)       # fp32 autograd of the reference vs the float64 adjoint
        rec.update({name + "/d_out": gw.numpy(), name + "/d_points": d_uv})
        print("triangulation", name, "backward: oracle adjoint == reference autograd (rel %.1e)" % gerr)
        rec.update({name + "/proj": P.numpy(), name + "/points": uv.numpy(), name + "/bk0": bk0, name + "/seed": np.int64(seed),
                    name + "/ref": ref.numpy(), name + "/svd": svd.numpy(), name + "/gt": X.numpy()})
        print("triangulation", name, "oracle == reference (rel %.1e); |sii - svd| max %.2e; |sii - gt| max %.2f"
              % (err, np.abs(ref.numpy() - svd.numpy()).max(), np.abs(ref.numpy() - X.numpy()).max()))
    np.savez_compressed(os.path.join(GOLD, "triangulation.npz"), **rec)


def glue_fixture():
    """the UNMODIFIED reference's HeatmapGenerator, flip_back (+ the flip-test merge of
    core/function.py:681-701), ToTensor + Normalize and GlobalAveragePoolingHead -> tests/golden/glue.npz; the oracle
    restatements (oracle/glue_oracle.py) are asserted equal."""
    import importlib.util
    from oracle import glue_oracle as G
    ref_shim.install()
    spec = importlib.util.spec_from_file_location(
        "ref_target_generators", os.path.join(ref_shim.REF_ROOT, "lib/dataset/target_generators/target_generators.py"))
    tg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tg)
    import utils.transforms as RT
    rec = {}
    g = torch.Generator().manual_seed(21)
    # ---- heat maps: 64 x 64, sigma 2 (RHD / MHP configs), 21 joints, 6 samples incl. edge cases ----
    J, res, sigma = 21, 64, 2
    joints = torch.rand(6, J, 3, generator=g) * torch.tensor([res + 8.0, res + 8.0, 1.0]) - torch.tensor([4.0, 4.0, 0.0])
    joints[..., 2] = (torch.rand(6, J, generator=g) < 0.85).float()
    joints[0, 0] = torch.tensor([0.0, 0.0, 1.0]); joints[0, 1] = torch.tensor([63.9, 63.9, 1.0])
    joints[0, 2] = torch.tensor([-0.5, 10.0, 1.0]); joints[0, 3] = torch.tensor([64.0, 10.0, 1.0])      # int(-0.5) = 0: inside; 64: outside
    joints[0, 4] = torch.tensor([3.2, 60.7, 1.0]); joints[0, 5] = torch.tensor([30.0, 30.0, 0.0])      # invisible
    gen = tg.HeatmapGenerator(res, J, sigma)
    hms = np.stack([gen(joints[b].numpy()) for b in range(6)])
    mine = np.stack([G.heatmap_generator(joints[b].numpy(), (res, res), sigma) for b in range(6)])
    assert np.array_equal(hms, mine)
    rec.update(hm_joints=joints.numpy(), hm_out=hms, hm_sigma=np.array(sigma), hm_res=np.array(res))
    # ---- flip_back + flip-test merge (RHD flip pairs: none are defined for hands in most configs; use a synthetic pairing) ----
    pairs = [[1, 2], [3, 4], [5, 8], [17, 20]]
    a = torch.randn(3, J, 16, 24, generator=g).numpy()
    bflip = torch.randn(3, J, 16, 24, generator=g).numpy()
    fb = RT.flip_back(bflip.copy(), pairs)
    assert np.array_equal(fb, G.flip_back(bflip, pairs))
    merged = {}
    for shift in (0, 1):
        t = torch.from_numpy(fb.copy())
        if shift:
            t[:, :, :, 1:] = t.clone()[:, :, :, 0:-1]                     # core/function.py:697-699
        merged[shift] = ((torch.from_numpy(a) + t) * 0.5).numpy()         # :701
        assert np.array_equal(merged[shift], G.flip_test_merge(a, bflip, pairs, bool(shift)))
    rec.update(flip_pairs=np.array(pairs), flip_a=a, flip_b=bflip, flip_back=fb, flip_merge0=merged[0], flip_merge1=merged[1])
    # ---- ToTensor + Normalize (the reference's own transform classes over torchvision) ----
    spec = importlib.util.spec_from_file_location(
        "ref_transforms", os.path.join(ref_shim.REF_ROOT, "lib/dataset/transforms/transforms.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    img = (torch.rand(2, 32, 48, 3, generator=g) * 256).clamp(0, 255).to(torch.uint8).numpy()
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    norm = []
    for b in range(2):
        t, _ = tr.ToTensor()(img[b], None)
        t, _ = tr.Normalize(mean=mean, std=std)(t, None)
        norm.append(t.numpy())
        assert np.allclose(G.normalize_u8(img[b], mean, std), norm[-1], rtol=0, atol=1e-6)
    rec.update(norm_img=img, norm_out=np.stack(norm), norm_mean=np.array(mean, np.float32), norm_std=np.array(std, np.float32))
    # ---- GlobalAveragePoolingHead (seeded default init; only outputs + checksums are stored) ----
    from models.pose_hrnet_volumetric import GlobalAveragePoolingHead
    torch.manual_seed(5)
    head = GlobalAveragePoolingHead(64, 32).eval()
    sd = head.state_dict()
    for k in sorted(sd):
        if k.endswith("running_mean"):
            sd[k].copy_(torch.randn(sd[k].shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            sd[k].copy_(torch.rand(sd[k].shape, generator=g) + 0.5)
    head.load_state_dict(sd)
    x = torch.randn(3, 64, 16, 16, generator=g)
    with torch.no_grad():
        out = head(x)
    mine = G.gap_head({"h." + k: v for k, v in head.state_dict().items()}, "h", x)
    assert torch.allclose(out, mine, rtol=1e-5, atol=1e-7)
    rec.update(gap_x=x.numpy(), gap_out=out.numpy(), gap_wsum=np.array(float(sum(v.double().abs().sum() for v in sd.values()))))
    for k in sorted(sd):
        if k.endswith(("running_mean", "running_var")):
            rec["gap_stat/" + k] = sd[k].numpy()
    # ---- cross-view fusion (row f2): the reference's Aggregation on 8 x 8 heat maps (12 FC layers of 64 x 64) ----
    from models.multiview_pose_hrnet import Aggregation
    torch.manual_seed(9)
    ag = Aggregation(ref_shim.to_attr({"MODEL": {"HEATMAP_SIZE": [8, 8]}})).eval()
    views = [torch.rand(2, 5, 8, 8, generator=g) for _ in range(4)]
    with torch.no_grad():
        fused = ag(views)
    mine = G.aggregation([m_.weight.weight.detach() for m_ in ag.aggre], views)
    for a_, b_ in zip(fused, mine):
        assert torch.allclose(a_, b_, rtol=1e-5, atol=1e-6)
    rec.update(agg_views=np.stack([v.numpy() for v in views]), agg_out=np.stack([f.numpy() for f in fused]),
               agg_wsum=np.array(float(sum(m_.weight.weight.double().abs().sum() for m_ in ag.aggre))),
               agg_keys=np.array(list(ag.state_dict().keys())))
    np.savez_compressed(os.path.join(GOLD, "glue.npz"), **rec)
    print("glue ok; oracle == reference (heat maps, flip_back / merge, normalisation, confidence head)")


def conditioned_fixtures():
    Y = "experiments/RHD/RHD_HRNet_w32_softmax_hm-pose2dloss_v1.yaml"
    YR = "experiments/RHD/RHD_HRNet_w32_max_hmloss_v1.yaml"
    train_conditioned_fixture("train_w32_softmax_contractive", Y, "softmax", contract=0.05)
    train_conditioned_fixture("train_w32_softmax_warm", Y, "softmax", warm_steps=WARM_STEPS)
    # BASELINE configs[4] geometry class: W48 raw variant + HeatmapLoss only
    train_conditioned_fixture("train_w48_raw_contractive", YR, "raw", B=4, contract=0.05, width_override=48)


if __name__ == "__main__":
    if "--triangulation-only" in sys.argv:
        triangulation_fixture()
        sys.exit(0)
    if "--glue-only" in sys.argv:
        glue_fixture()
        sys.exit(0)
    if "--conditioned-only" in sys.argv:
        conditioned_fixtures()
        sys.exit(0)
    if "--train-only" in sys.argv:
        train_fixture("train_w32_softmax", "experiments/RHD/RHD_HRNet_w32_softmax_hm-pose2dloss_v1.yaml", "softmax",
                      trainable_temp=True)
        train_fixture("train_w32_raw", "experiments/RHD/RHD_HRNet_w32_max_hmloss_v1.yaml", "raw")
        sys.exit(0)
    os.makedirs(GOLD, exist_ok=True)
    decode_fixture()
    loss_fixture()
    Y = "experiments/RHD/RHD_HRNet_w32_softmax_hm-pose2dloss_v1.yaml"
    YR = "experiments/RHD/RHD_HRNet_w32_max_hmloss_v1.yaml"
    net_fixture("hrnet_w32_softmax_default", Y, "softmax", perturb=False)   # the spec's "shared random-init weights"
    net_fixture("hrnet_w32_softmax", Y, "softmax")
    net_fixture("hrnet_w32_softmax_sharp", Y, "softmax", sharpen=True)
    net_fixture("hrnet_w32_raw", YR, "raw")
    net_fixture("hrnet_w48_softmax_rect", "experiments/RHD/RHD_HRNet_w48_trainable_softmax_hm-pose2dloss_v1.yaml",
                "softmax", H=128, W=96, B=2)
    train_fixture("train_w32_softmax", Y, "softmax", trainable_temp=True)
    train_fixture("train_w32_raw", YR, "raw")
    conditioned_fixtures()
    glue_fixture()
    triangulation_fixture()
