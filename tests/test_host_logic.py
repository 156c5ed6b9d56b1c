"""CPU: host-side logic - cfg handling, module tree, error behaviour, state_dict interchange with the reference."""
import pytest
import torch

from hrnet_b200 import arch as A
from hrnet_b200.config import make_cfg, to_cfg
from hrnet_b200.models import pose_hrnet, pose_hrnet_softmax


def test_cfg_dual_access_and_arch():
    cfg = make_cfg(48)
    assert cfg.MODEL.EXTRA.STAGE4.NUM_CHANNELS == cfg["MODEL"]["EXTRA"]["STAGE4"]["NUM_CHANNELS"] == [48, 96, 192, 384]
    a = A.arch_from_cfg(cfg)
    assert a.head_channels == 720 and a.modules == (1, 4, 3)
    convs = [s for s in A.layer_specs(a) if isinstance(s, A.Conv)]
    bns = [s for s in A.layer_specs(a) if isinstance(s, A.BN)]
    assert (len(convs), len(bns)) == (307, 306)           # SURVEY §6 census


def test_branch_mismatch_raises_value_error_like_reference():
    cfg = make_cfg(32)
    cfg.MODEL.EXTRA.STAGE3.NUM_BLOCKS = [4, 4]
    with pytest.raises(ValueError, match="NUM_BRANCHES"):
        pose_hrnet.get_pose_net(cfg, is_train=False)
    cfg = make_cfg(32)
    cfg.MODEL.EXTRA.STAGE2.NUM_CHANNELS = [32]
    with pytest.raises(ValueError, match="NUM_CHANNELS"):
        pose_hrnet.get_pose_net(cfg, is_train=False)


def test_missing_pretrained_file_raises_value_error():
    cfg = make_cfg(32, init_weights=True)
    cfg.MODEL.PRETRAINED = "/nonexistent/model.pth"
    with pytest.raises(ValueError, match="does not exist"):
        pose_hrnet.get_pose_net(cfg, is_train=True)


def test_module_tree_surface():
    m = pose_hrnet_softmax.get_pose_net(make_cfg(32, trainable_softmax=True), is_train=False)
    assert m.trainable_temp.requires_grad and float(m.trainable_temp) == 1.0
    assert m.transition2[0] is None and m.transition2[2] is not None and len(m.transition3) == 4
    assert len(m.stage4) == 3 and len(m.stage4[0].branches) == 4 and len(m.layer1) == 4
    assert m.last_layer[0].bias is not None and m.last_layer[3].out_channels == 21
    assert sum(p.numel() for p in m.stage4.parameters()) > 0        # wrappers freeze/unfreeze by sub-tree
    m2 = pose_hrnet.get_pose_net(make_cfg(32, softmax=False), is_train=False)
    assert "trainable_temp" not in m2.state_dict() and len(m2.state_dict()) == 1839


def test_init_weights_statistics():
    m = pose_hrnet.get_pose_net(make_cfg(32, softmax=False, init_weights=True), is_train=True)
    w = m.stage3[1].branches[2][0].conv1.weight
    assert abs(float(w.std()) - 0.001) < 2e-4 and float(m.last_layer[3].bias.abs().sum()) == 0.0


def test_no_cpu_fallback():
    m = pose_hrnet_softmax.get_pose_net(make_cfg(32), is_train=False).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 256, 256))
    m.train()                                   # training runs on the CUDA engine too: CPU tensors / modules are refused
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 256, 256))
    from hrnet_b200.core.loss import HeatmapLoss
    from hrnet_b200.utils.heatmap_decoding import get_final_preds
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HeatmapLoss()(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        get_final_preds(torch.zeros(1, 2, 4, 4))
    with pytest.raises(AssertionError):
        get_final_preds(torch.zeros(2, 4, 4).numpy())


def test_state_dict_interchange_with_reference():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    ref_hrnet, ref_softmax, *_ = ref_shim.modules()
    cfg = ref_shim.load_cfg()
    ref = ref_softmax.get_pose_net(cfg, is_train=False)
    ours = pose_hrnet_softmax.get_pose_net(cfg, is_train=False)          # the reference's own cfg object
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    ours.load_state_dict({"module." + k: v for k, v in ref.state_dict().items()}, strict=False)  # DP prefix: no match, no crash
    for k, v in ref.state_dict().items():
        assert ours.state_dict()[k].shape == v.shape and ours.state_dict()[k].dtype == v.dtype


def test_product_synthetic_data_equals_the_test_fixtures():
    """bench.py's product arm draws its workload from hrnet_b200.synthetic (it must not import oracle/): same tensors as
    the fixtures the goldens and the CPU baseline use"""
    import torch
    from hrnet_b200 import synthetic
    from oracle import fixtures
    assert torch.equal(synthetic.images(2, 64, 96, seed=5), fixtures.images(2, 64, 96, seed=5))
    for a, b in zip(synthetic.targets(3, 21, 16, 24, seed=7), fixtures.targets(3, 21, 16, 24, seed=7)):
        assert torch.equal(a, b)


def test_triangulation_mirror_host_logic():
    """row (f) drop-in: same start-vector stream as the reference's per-joint loop, no CPU fallback"""
    import numpy as np
    import pytest
    import torch
    from hrnet_b200.utils import misc
    from oracle import triangulation_oracle as T
    torch.manual_seed(5)
    mine = misc._start_vectors(3, 4, "cpu").numpy()
    assert np.array_equal(mine, T.start_vectors(3, 4, 5))
    with pytest.raises(RuntimeError, match="CUDA"):
        misc.DLT_sii_pytorch(torch.zeros(2, 4, 2), torch.zeros(2, 4, 3, 4))
    h = torch.tensor([[2.0, 4.0, 6.0, 2.0]])
    assert torch.equal(misc.homogeneous_to_euclidean(h), torch.tensor([[1.0, 2.0, 3.0]]))


def test_fused_statistics_eligibility_rule():
    """which conv launches may also reduce the BatchNorm statistics (include/hrnb.h stats_sums): one N tile of 16 / 32 / 64
    channels on the flat-shift PF8 path, no residual, no ReLU"""
    from hrnet_b200 import _lib
    from hrnet_b200.ops import stats_eligible

    def params(cout, bn, flags=0, res=None):
        p = _lib.ConvParams()
        p.cout, p.BN, p.flags, p.res = cout, bn, flags, res
        return p
    assert stats_eligible(params(32, 32)) and stats_eligible(params(64, 64)) and stats_eligible(params(16, 16))
    assert stats_eligible(params(64, 64, _lib.HRNB_CONV_IN_PHASES))          # stride 2 over phase-split input: flat-shift
    assert not stats_eligible(params(64, 32))                                 # two N tiles
    assert not stats_eligible(params(128, 128)) and not stats_eligible(params(48, 48))
    assert not stats_eligible(params(32, 32, _lib.HRNB_CONV_RELU))
    assert not stats_eligible(params(32, 32, _lib.HRNB_CONV_GATHER))
    assert not stats_eligible(params(32, 32, _lib.HRNB_CONV_OUT_NCHW))
    assert not stats_eligible(params(32, 32, 0, res=0x1000))


def test_named_parameter_and_module_order_match_the_reference():
    """an index-keyed optimizer.state_dict() of a reference checkpoint (tools/train.py:285,380) must line up
    with our parameters, and init_weights' walk over modules() must draw the RNG in the reference's order"""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    ref_hrnet, ref_softmax, *_ = ref_shim.modules()
    for mod_ref, mod_ours, yaml_rel in ((ref_softmax, pose_hrnet_softmax, "experiments/RHD/RHD_HRNet_w32_softmax_hm-pose2dloss_v1.yaml"),
                                        (ref_hrnet, pose_hrnet, "experiments/RHD/RHD_HRNet_w32_max_hmloss_v1.yaml")):
        cfg = ref_shim.load_cfg(yaml_rel)
        ref = mod_ref.get_pose_net(cfg, is_train=False)
        ours = mod_ours.get_pose_net(cfg, is_train=False)
        assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
        assert [n for n, _ in ours.named_buffers()] == [n for n, _ in ref.named_buffers()]
        assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
        # init_weights (normal std 0.001, BN constants) under the same seed gives the same tensors
        torch.manual_seed(3); ref.init_weights("")
        torch.manual_seed(3); ours.init_weights("")
        for (k, a), (_, b) in zip(ours.state_dict().items(), ref.state_dict().items()):
            assert torch.equal(a, b), k


def test_unsupported_widths_fail_early_with_a_clear_message():
    from hrnet_b200.config import make_cfg
    with pytest.raises(ValueError, match="multiple of 16"):
        pose_hrnet.get_pose_net(make_cfg(18, softmax=False), is_train=False)


def test_grad_allreduce_bucket_bounds_and_cpu_form():
    from hrnet_b200.parallel import GradAllReduce
    ar = GradAllReduce(1000, n_buckets=3)
    assert ar.bounds[0][0] == 0 and ar.bounds[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(ar.bounds[:-1], ar.bounds[1:]))
    assert not ar.overlap and ar.mean_scale == 1.0
    t = torch.arange(1000, dtype=torch.float32)
    assert torch.equal(ar(t.clone()), t)          # world 1: identity


def test_volumetric_backbone_mirrors_the_reference_module_tree():
    """row f1: pose_hrnet_volumetric.get_pose_net - same state_dict keys, parameter order and seeded init as the reference
    (VOL_CONFIDENCES head created between stage4 and last_layer)"""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    ref_shim.install()
    from models import pose_hrnet_volumetric as RV
    from hrnet_b200.models import pose_hrnet_volumetric as OV
    cfg = ref_shim.load_cfg()
    cfg.MODEL["ALG_CONFIDENCES"], cfg.MODEL["VOL_CONFIDENCES"], cfg.MODEL["TRAINABLE_SOFTMAX"] = False, True, True
    torch.manual_seed(0)
    ref = RV.get_pose_net(cfg, is_train=False)
    torch.manual_seed(0)
    ours = OV.get_pose_net(cfg, is_train=False)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    for (k, a), (_, b) in zip(ours.state_dict().items(), ref.state_dict().items()):
        assert torch.equal(a, b), k
    assert all(not n.startswith("vol_confidences") for n, _ in ours.engine_parameters())
    assert len(ours.engine_parameters()) == len(list(ours.named_parameters())) - len(list(ours.vol_confidences.parameters()))
    # the reference cannot even construct with ALG_CONFIDENCES: true (NameError at pose_hrnet_volumetric.py:375)
    cfg.MODEL["ALG_CONFIDENCES"] = True
    with pytest.raises(NameError):
        RV.get_pose_net(cfg, is_train=False)


def test_algebraic_triangulation_net_host_logic():
    from hrnet_b200.models.triangulation import AlgebraicTriangulationNet
    from hrnet_b200.config import make_cfg
    cfg = make_cfg(32, trainable_softmax=True)
    cfg.MODEL["BACKBONE_NAME"], cfg.MODEL["ALG_CONFIDENCES"], cfg.MODEL["BACKBONE_MODEL_PATH"] = "pose_hrnet_volumetric", False, ""
    net = AlgebraicTriangulationNet(cfg)
    free = {n.split(".")[0] for n, p in net.backbone.named_parameters() if p.requires_grad}
    assert free == {"stage4", "last_layer"}                  # lib/models/triangulation.py:205-215
    cfg.MODEL["ALG_CONFIDENCES"] = True
    net2 = AlgebraicTriangulationNet(cfg)
    assert hasattr(net2.backbone, "alg_confidences")
    with pytest.raises(RuntimeError, match="ALG_CONFIDENCES"):
        net2(torch.zeros(1, 4, 3, 64, 64), torch.zeros(1, 4, 3, 4))
    cfg.MODEL["BACKBONE_NAME"] = "pose_resnet"
    with pytest.raises(ValueError, match="BACKBONE_NAME"):
        AlgebraicTriangulationNet(cfg)
