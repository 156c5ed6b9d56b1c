"""CPU: the oracle restatements against the golden vectors generated from the unmodified reference
(oracle/make_golden.py), and the product's seeded initialisation / key layout against the same."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle, fixtures, hrnet_oracle, loss_oracle
from hrnet_b200.config import make_cfg
from hrnet_b200.models import pose_hrnet, pose_hrnet_softmax


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _key_hash(sd):
    return hashlib.sha256("\n".join(sorted("%s %s %s" % (k, tuple(v.shape), v.dtype) for k, v in sd.items())).encode()).hexdigest()


def _product_model(width, variant, trainable=False):
    cfg = make_cfg(width, softmax=(variant == "softmax"), trainable_softmax=trainable)
    mod = pose_hrnet_softmax if variant == "softmax" else pose_hrnet
    torch.manual_seed(0)
    return mod.get_pose_net(cfg, is_train=False), cfg


NETS = [("hrnet_w32_softmax_default", 32, "softmax", False), ("hrnet_w32_softmax", 32, "softmax", False),
        ("hrnet_w32_softmax_sharp", 32, "softmax", True),
        ("hrnet_w32_raw", 32, "raw", False), ("hrnet_w48_softmax_rect", 48, "softmax", False)]


@pytest.mark.parametrize("name,width,variant,sharp", NETS)
def test_network_oracle_and_seeded_init_match_reference(golden_dir, name, width, variant, sharp):
    g = _load(golden_dir, name + ".npz")
    model, cfg = _product_model(width, variant)
    sd = model.state_dict()
    # same keys / shapes / dtypes as the reference module tree
    assert len(sd) == int(g["n_keys"])
    assert _key_hash(sd) == str(g["key_list_hash"])
    # same seeded default initialisation (parameter creation order consumes the RNG identically)
    for k, s in zip(g["init_keys"], g["init_sums"]):
        assert np.isclose(float(sd[str(k)].double().abs().sum()), float(s), rtol=1e-9), k
    total = float(sum(v.double().abs().sum() for v in sd.values() if v.dtype.is_floating_point))
    assert np.isclose(total, float(g["init_total_abs"]), rtol=1e-9)
    # oracle forward on the perturbed weights == reference forward
    perturb = not name.endswith("_default")
    if perturb:
        fixtures.perturb_state_dict(sd)
    if sharp:
        fixtures.sharpen_head(sd)
    if variant == "softmax" and perturb:
        sd["trainable_temp"].fill_(1.7)
    x = fixtures.images(int(g["B"]), int(g["H"]), int(g["W"]))
    arch = hrnet_oracle.Arch.from_cfg(cfg)
    out = hrnet_oracle.forward(sd, x, arch, variant)
    if variant == "softmax":
        heat, feat, temp, logits = out
        assert np.allclose(heat[:, :, ::2, ::2].numpy(), g["heat"], rtol=2e-4, atol=1e-9)
        assert np.allclose(logits[:, :, ::2, ::2].numpy(), g["logits"], rtol=1e-3, atol=1e-5)
        assert np.allclose(feat[:, ::16, ::4, ::4].numpy(), g["feat"], rtol=1e-3, atol=1e-5)
        assert np.isclose(float(heat.double().sum()), float(g["heat_sum"]), rtol=1e-6)
        assert np.allclose(decode_oracle.spatial_expectation2d(heat.numpy()), g["soft_coords"], atol=2e-3)
        p, mv = decode_oracle.get_max_preds(heat.numpy())
        # argmax of nearly flat default-init maps may flip under thread-count dependent summation order;
        # the sharpened set must agree exactly
        if sharp:
            assert np.array_equal(p, g["max_preds"])
        assert np.allclose(mv, g["maxvals"], rtol=1e-3)
    else:
        logits, feat = out
        assert np.allclose(logits[:, :, ::2, ::2].numpy(), g["logits"], rtol=1e-3, atol=1e-5)
        assert np.allclose(feat[:, ::4, ::4, ::4].numpy(), g["feat"], rtol=1e-3, atol=1e-5)


def test_decode_oracle_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "decode.npz")
    for tag in ("sq", "rect", "j20"):
        for nm in ("soft", "edge"):
            a = g["%s_%s_in" % (tag, nm)]
            p, mv = decode_oracle.get_max_preds(a.copy())
            assert np.array_equal(p, g["%s_%s_maxpreds" % (tag, nm)])
            assert np.array_equal(mv, g["%s_%s_maxvals" % (tag, nm)])
            assert np.array_equal(decode_oracle.argmax_hstride(a), g["%s_%s_hstride" % (tag, nm)])
            assert np.allclose(decode_oracle.spatial_expectation2d(a), g["%s_%s_expect" % (tag, nm)], rtol=1e-4, atol=1e-3)
            for pp in (0, 1):
                fp, _ = decode_oracle.final_preds(a.copy(), g[tag + "_center"], g[tag + "_scale"], bool(pp))
                assert np.allclose(fp, g["%s_%s_final%d" % (tag, nm, pp)], rtol=1e-5, atol=1e-4)
        assert np.allclose(decode_oracle.spatial_softmax(g[tag + "_logits"], 1.7), g[tag + "_softmax17"], rtol=1e-5, atol=1e-8)


def test_decode_edge_semantics(golden_dir):
    """The reference's documented corner cases: all-zero, all-negative, ties, H-stride bug."""
    g = _load(golden_dir, "decode.npz")
    a = g["rect_edge_in"]                                  # 24 x 16 maps
    p, mv = decode_oracle.get_max_preds(a.copy())
    assert (p[0, 0] == 0).all() and mv[0, 0, 0] == 0       # all-zero: masked to (0, 0)
    assert (p[0, 1] == 0).all() and mv[0, 1, 0] < 0        # all-negative: coords zeroed, maxval kept
    assert tuple(p[0, 2]) == (5.0, 3.0)                    # tie: first index in row-major order
    hs = decode_oracle.argmax_hstride(a)
    idx = a.reshape(a.shape[0], a.shape[1], -1).argmax(2)
    assert np.array_equal(hs[..., 0], idx % 24) and np.array_equal(hs[..., 1], idx // 24)   # H used as stride


def test_loss_oracle_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "loss.npz")
    for tag in ("a", "b"):
        for mode in ("l2", "l1"):
            assert np.isclose(loss_oracle.heatmap_loss(g[tag + "_pred"], g[tag + "_gt"], mode), g["%s_hm_%s" % (tag, mode)], rtol=1e-5)
            assert np.allclose(loss_oracle.heatmap_loss_grad(g[tag + "_pred"], g[tag + "_gt"], mode),
                               g["%s_hm_%s_grad" % (tag, mode)], rtol=1e-5, atol=1e-8)
        vis = g[tag + "_vis"]
        for vtag, v in (("vis", vis), ("novis", None), ("zerovis", np.zeros_like(vis))):
            assert np.isclose(loss_oracle.pose2d_loss(g[tag + "_pp"], g[tag + "_xy"], v), g["%s_p2d_%s" % (tag, vtag)], rtol=1e-5)
            assert np.allclose(loss_oracle.pose2d_loss_grad(g[tag + "_pp"], g[tag + "_xy"], v),
                               g["%s_p2d_%s_grad" % (tag, vtag)], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("name,variant", [("train_w32_softmax", "softmax"), ("train_w32_raw", "raw")])
def test_train_oracle_reproduces_reference_training_step(golden_dir, name, variant):
    """oracle/train_oracle.py (torch-CPU autograd restatement of one reference training step) against the losses,
    gradients, running statistics and Adam-updated parameters the UNMODIFIED reference produced
    (oracle/make_golden.py train_fixture)."""
    import numpy as np
    import torch
    from oracle import fixtures, hrnet_oracle, train_oracle
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet, pose_hrnet_softmax
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, H, W = int(g["B"]), int(g["H"]), int(g["W"])
    cfg = make_cfg(32, softmax=(variant == "softmax"), trainable_softmax=bool(g["trainable_temp"]))
    torch.manual_seed(0)
    m = (pose_hrnet_softmax if variant == "softmax" else pose_hrnet).get_pose_net(cfg, is_train=False)
    sd = m.state_dict()
    fixtures.perturb_state_dict(sd)
    x = fixtures.images(B, H, W)
    gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4)
    o = train_oracle.train_step(sd, x, gt, xy, vis, hrnet_oracle.Arch.from_cfg(cfg), variant,
                                trainable_temp=bool(g["trainable_temp"]))
    assert np.allclose(o["losses"], g["losses"], rtol=1e-4)
    gmax = max(float(np.abs(g["grad/" + str(k)]).max()) for k in g["keys"])
    for k in g["keys"]:
        k = str(k)
        ref = g["grad/" + k]
        got = fixtures.sample(o["grads"][k]).numpy()
        # floor: gradients that are mathematically zero (a bias in front of BatchNorm / softmax) are rounding noise
        assert np.abs(got - ref).max() <= 2e-3 * max(np.abs(ref).max(), 1e-4 * gmax), k
        if np.abs(ref).max() > 1e-4 * gmax:
            assert np.isclose(float(o["grads"][k].double().norm()), float(g["gnorm/" + k]), rtol=1e-3)
        d = np.abs(fixtures.sample(o["state"][k]).numpy() - g["after/" + k])
        assert d.max() <= 2.1e-3 and (d > 1e-5).mean() < 0.02, k          # first Adam step: +-lr per element
    for k in ("bn1", "stage3.0.branches.1.2.bn1", "last_layer.1"):
        for s in (".running_mean", ".running_var"):
            assert np.allclose(o["state"][k + s].numpy(), g["after/" + k + s], rtol=1e-4, atol=1e-6), k + s


def test_triangulation_oracle_matches_reference_golden(golden_dir):
    """oracle/triangulation_oracle.py against the values of the unmodified reference DLT_sii_pytorch
    (lib/utils/misc.py:64-97, called per joint like lib/models/triangulation.py:258-261) stored by oracle/make_golden.py;
    the start vectors are re-drawn from the recorded seed; edge cases: two views, 8 views, J = 20."""
    from oracle import triangulation_oracle as T
    g = _load(golden_dir, "triangulation.npz")
    for case in ("mhp4", "two_views", "eight_views_j20"):
        P, uv, ref = g[case + "/proj"], g[case + "/points"], g[case + "/ref"]
        B, V, J, _ = uv.shape
        bk0 = T.start_vectors(B, J, int(g[case + "/seed"]))
        assert np.array_equal(bk0, g[case + "/bk0"])
        assert np.allclose(np.linalg.norm(bk0, axis=-1), 1.0, atol=1e-6)
        out = T.triangulate_joints(uv, P, bk0)
        scale = np.abs(ref).max()
        assert np.abs(out - ref).max() < 2e-5 * scale, case
        # two inverse iterations land on the exact smallest singular vector up to fp32 conditioning, and near the truth
        svd = np.stack([T.svd_triangulation(uv[:, :, k], P) for k in range(J)], axis=1)
        assert np.abs(svd - g[case + "/svd"]).max() < 2e-2
        assert np.abs(out - svd).max() < 0.5 and np.abs(out - g[case + "/gt"]).max() < 10.0


@pytest.mark.parametrize("name,variant", [("train_w32_softmax_contractive", "softmax"), ("train_w48_raw_contractive", "raw")])
def test_train_oracle_reproduces_conditioned_reference_goldens(golden_dir, name, variant):
    """the well-conditioned training goldens (oracle/make_golden.py train_conditioned_fixture, generated by the UNMODIFIED
    reference): the oracle port reproduces their losses and (sampled) gradients from the seeded + contracted weights"""
    from oracle import hrnet_oracle, train_oracle
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, H, W, width = int(g["B"]), int(g["H"]), int(g["W"]), int(g["width"])
    cfg = make_cfg(width, softmax=(variant == "softmax"), trainable_softmax=True, image_size=(H, W))
    torch.manual_seed(0)
    sd = (pose_hrnet_softmax if variant == "softmax" else pose_hrnet).get_pose_net(cfg, is_train=False).state_dict()
    fixtures.perturb_state_dict(sd)
    fixtures.contract_state_dict(sd, float(g["contract"]))
    for k in g["keys"]:
        assert np.allclose(fixtures.sample(sd[str(k)]).numpy(), g["weight/" + str(k)], rtol=1e-6, atol=1e-8), k
    x = fixtures.images(B, H, W)
    gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4)
    o = train_oracle.train_step(sd, x, gt, xy, vis, hrnet_oracle.Arch.from_cfg(cfg), variant, trainable_temp=True, adam=False)
    assert np.allclose(o["losses"], g["losses"], rtol=1e-4)
    gmax = float(g["gmax_all"].max())
    norms = dict(zip([str(k) for k in g["all_keys"]], g["gnorm_all"]))
    for k in g["keys"]:
        k = str(k)
        ref = g["grad/" + k]
        assert np.abs(fixtures.sample(o["grads"][k]).numpy() - ref).max() <= 2e-3 * max(np.abs(ref).max(), 1e-4 * gmax), k
    for k, v in o["grads"].items():
        if float(v.abs().max()) > 1e-4 * gmax:
            assert np.isclose(float(v.double().norm()), norms[k], rtol=2e-3), k


def test_warm_golden_records_the_reference_trajectory(golden_dir):
    """train_w32_softmax_warm.npz: 60-step loss trajectory of the unmodified reference + the oracle port's deviation from
    it (chaos floor, recorded by make_golden) + step 0 reproduced here by the oracle port"""
    from oracle import hrnet_oracle, train_oracle
    from oracle.make_golden import warm_batch
    g = np.load(os.path.join(golden_dir, "train_w32_softmax_warm.npz"))
    assert g["trajectory"].shape == (int(g["warm_steps"]), 3) and int(g["warm_steps"]) == 60
    assert 0 < float(g["oracle_vs_ref_traj_dev"]) < 0.15 and float(g["oracle_vs_ref_worst_cos"]) > 0.99999
    B, H, W = int(g["B"]), int(g["H"]), int(g["W"])
    cfg = make_cfg(32, softmax=True, trainable_softmax=True, image_size=(H, W))
    torch.manual_seed(0)
    sd = pose_hrnet_softmax.get_pose_net(cfg, is_train=False).state_dict()
    fixtures.perturb_state_dict(sd)
    o = train_oracle.train_step(sd, *warm_batch(0, B, H, W), hrnet_oracle.Arch.from_cfg(cfg), "softmax", trainable_temp=True)
    assert np.allclose(o["losses"], g["trajectory"][0], rtol=1e-4)
    assert o["opt_state"] is not None and int(o["state"]["bn1.num_batches_tracked"]) == 1


def test_triangulation_adjoint_matches_reference_autograd(golden_dir):
    """oracle/triangulation_oracle.dlt_sii_backward (the chain the CUDA adjoint kernel follows) against the gradients the
    unmodified reference's autograd produced through DLT_sii_pytorch"""
    from oracle import triangulation_oracle as T
    g = np.load(os.path.join(golden_dir, "triangulation.npz"))
    for case in ("mhp4", "two_views", "eight_views_j20"):
        P, uv, bk0, d_out, ref = (g[case + "/" + k] for k in ("proj", "points", "bk0", "d_out", "d_points"))
        mine = np.stack([T.dlt_sii_backward(uv[:, :, k], P, bk0[k], d_out[:, k]) for k in range(uv.shape[2])], axis=2)
        assert np.abs(mine - ref).max() < 2e-3 * np.abs(ref).max(), case


def test_glue_oracle_matches_reference_golden(golden_dir):
    """oracle/glue_oracle.py against the outputs of the unmodified reference's HeatmapGenerator, flip_back (+ flip-test merge)
    and ToTensor + Normalize stored in tests/golden/glue.npz"""
    from oracle import glue_oracle as G
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    res, sigma = int(g["hm_res"]), int(g["hm_sigma"])
    for b in range(g["hm_joints"].shape[0]):
        assert np.array_equal(G.heatmap_generator(g["hm_joints"][b], (res, res), sigma), g["hm_out"][b])
    assert g["hm_out"][0, 2].max() == 1.0 and g["hm_out"][0, 3].max() == 0.0 and g["hm_out"][0, 5].max() == 0.0   # edge semantics
    pairs = g["flip_pairs"].tolist()
    assert np.array_equal(G.flip_back(g["flip_b"], pairs), g["flip_back"])
    assert np.array_equal(G.flip_test_merge(g["flip_a"], g["flip_b"], pairs, False), g["flip_merge0"])
    assert np.array_equal(G.flip_test_merge(g["flip_a"], g["flip_b"], pairs, True), g["flip_merge1"])
    for b in range(2):
        assert np.allclose(G.normalize_u8(g["norm_img"][b], g["norm_mean"], g["norm_std"]), g["norm_out"][b], atol=1e-6)


def test_confidence_head_seeded_init_matches_reference(golden_dir):
    """the GlobalAveragePoolingHead mirror draws the reference's seeded weights (same construction order) and the oracle
    restatement reproduces the reference's output from them"""
    from oracle import glue_oracle as G
    from hrnet_b200.models.pose_hrnet_volumetric import GlobalAveragePoolingHead
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    torch.manual_seed(5)
    head = GlobalAveragePoolingHead(64, 32)
    sd = head.state_dict()
    for k in sd:
        if k.endswith(("running_mean", "running_var")):
            sd[k].copy_(torch.from_numpy(g["gap_stat/" + k]))
    assert np.isclose(float(sum(v.double().abs().sum() for v in sd.values())), float(g["gap_wsum"]), rtol=1e-9)
    out = G.gap_head({"h." + k: v for k, v in sd.items()}, "h", torch.from_numpy(g["gap_x"]))
    assert np.allclose(out.numpy(), g["gap_out"], rtol=1e-5, atol=1e-7)


def test_cross_view_aggregation_oracle_and_seeded_init_match_reference(golden_dir):
    """row f2: the Aggregation mirror has the reference's state-dict keys and seeded weights; the oracle restatement reproduces
    the reference's fused views from them"""
    from oracle import glue_oracle as G
    from hrnet_b200.models.multiview_pose_hrnet import Aggregation
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    torch.manual_seed(9)
    ag = Aggregation({"MODEL": {"HEATMAP_SIZE": [8, 8]}})
    assert list(ag.state_dict().keys()) == [str(k) for k in g["agg_keys"]]
    assert np.isclose(float(sum(m.weight.weight.double().abs().sum() for m in ag.aggre)), float(g["agg_wsum"]), rtol=1e-9)
    views = [torch.from_numpy(v) for v in g["agg_views"]]
    outs = G.aggregation([m.weight.weight.detach() for m in ag.aggre], views)
    for a, b in zip(outs, g["agg_out"]):
        assert np.allclose(a.numpy(), b, rtol=1e-5, atol=1e-6)
