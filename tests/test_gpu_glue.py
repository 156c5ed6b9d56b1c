"""-m gpu: the loop-glue kernels through their drop-in mirrors against tests/golden/glue.npz, which
holds the outputs of the UNMODIFIED reference classes (oracle/make_golden.py glue_fixture)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "glue.npz"))


def test_heatmap_generator_matches_reference(g):
    from hrnet_b200.dataset.target_generators import HeatmapGenerator
    joints = torch.from_numpy(g["hm_joints"]).cuda()
    gen = HeatmapGenerator(int(g["hm_res"]), joints.shape[1], float(g["hm_sigma"]))
    out = gen(joints).cpu().numpy()
    ref = g["hm_out"]
    assert out.shape == ref.shape
    assert np.array_equal(out == 0, ref == 0)                      # the same support (patch extent, visibility, bounds)
    assert np.abs(out - ref).max() <= 1.2e-7                       # exp in float64 on both sides; <= 1 float32 ulp
    one = gen(joints[2]).cpu().numpy()                             # the reference's per-sample call form [J, 3]
    assert np.array_equal(one, out[2])
    rect = HeatmapGenerator((32, 48), joints.shape[1], 1.5)(joints[:2])     # non-square / non-integer sigma: runs, peak <= 1
    assert rect.shape == (2, joints.shape[1], 32, 48) and float(rect.max()) <= 1.0
    with pytest.raises(RuntimeError, match="CUDA"):
        gen(joints.cpu())


def test_heatmap_generator_against_oracle_on_rectangular_maps(g):
    from oracle import glue_oracle as G
    from hrnet_b200.dataset.target_generators import HeatmapGenerator
    gen_ = torch.Generator().manual_seed(3)
    joints = torch.rand(4, 20, 3, generator=gen_) * torch.tensor([80.0, 100.0, 1.0]) - torch.tensor([4.0, 4.0, 0.3])
    out = HeatmapGenerator((96, 72), 20, 2)(joints.cuda()).cpu().numpy()
    ref = np.stack([G.heatmap_generator(joints[b].numpy(), (96, 72), 2) for b in range(4)])
    assert np.array_equal(out == 0, ref == 0) and np.abs(out - ref).max() <= 1.2e-7


def test_flip_back_and_flip_test_merge_match_reference(g):
    from hrnet_b200.utils.transforms import flip_back, flip_test_merge
    pairs = g["flip_pairs"].tolist()
    a, b = torch.from_numpy(g["flip_a"]).cuda(), torch.from_numpy(g["flip_b"]).cuda()
    assert np.array_equal(flip_back(b, pairs).cpu().numpy(), g["flip_back"])
    assert np.array_equal(flip_test_merge(a, b, pairs, False).cpu().numpy(), g["flip_merge0"])
    assert np.array_equal(flip_test_merge(a, b, pairs, True).cpu().numpy(), g["flip_merge1"])
    with pytest.raises(AssertionError):
        flip_back(b[0], pairs)


def test_normalisation_folded_into_the_stem(g):
    """uint8 NHWC images through forward_images == fp32 normalised NCHW images through forward (same kernels after the stem's
    im2col; the im2col slab itself must be bit-identical)"""
    from hrnet_b200 import _lib
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet_softmax
    from hrnet_b200.ops import PF8, stem_im2col
    import ctypes as C
    img = torch.from_numpy(g["norm_img"]).cuda()                                     # [2, 32, 48, 3] uint8
    ref = torch.from_numpy(g["norm_out"]).cuda()                                     # reference ToTensor + Normalize, [2, 3, 32, 48]
    B, H, W, _ = img.shape
    a, b = PF8(B, 32, H // 2, W // 2), PF8(B, 32, H // 2, W // 2)
    stem_im2col(ref, a)
    mean, std = (C.c_float * 3)(*g["norm_mean"].tolist()), (C.c_float * 3)(*g["norm_std"].tolist())
    _lib.check(_lib.lib().hrnb_stem_im2col_u8(img.data_ptr(), mean, std, b.ptr, b.ps, B, H, W, _lib.stream_ptr()))
    torch.cuda.synchronize()
    d = (a.to_nchw() - b.to_nchw()).abs().max().item()
    assert d <= 2 ** -7 * 3.0, d            # both sides round the same fp32 value to bf16; the golden is within 1e-6 of ours
    assert b.padding_is_zero()
    # whole network: 64 x 64 crop replicated to a legal input size
    torch.manual_seed(0)
    m = pose_hrnet_softmax.get_pose_net(make_cfg(32, image_size=(64, 64)), is_train=False).cuda().eval()
    big = torch.randint(0, 256, (2, 64, 64, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).cuda()
    x = ((big.permute(0, 3, 1, 2).float() / 255) - torch.tensor(m.IMAGENET_MEAN).view(1, 3, 1, 1).cuda()) / torch.tensor(m.IMAGENET_STD).view(1, 3, 1, 1).cuda()
    h_ref = m(x)[0]
    h_u8 = m.forward_images(big)[0]
    assert (h_ref - h_u8).abs().max().item() <= 2e-2 * h_ref.abs().max().item()


def test_confidence_head_matches_reference(g):
    from oracle import glue_oracle as G
    from hrnet_b200.models.pose_hrnet_volumetric import GlobalAveragePoolingHead
    torch.manual_seed(5)
    head = GlobalAveragePoolingHead(64, 32).eval()
    sd = head.state_dict()
    for k in sd:                                   # seeded default init (same RNG stream as the reference's constructor) + stored statistics
        if k.endswith(("running_mean", "running_var")):
            sd[k].copy_(torch.from_numpy(g["gap_stat/" + k]))
    head.load_state_dict(sd)
    x = torch.from_numpy(g["gap_x"])
    ref = G.gap_head({"h." + k: v for k, v in head.state_dict().items()}, "h", x)
    assert torch.allclose(ref, torch.from_numpy(g["gap_out"]), rtol=1e-5, atol=1e-7)     # oracle on our seeded weights == the reference's output
    out = head.cuda()(x.cuda()).cpu()
    assert out.shape == ref.shape
    assert (out - torch.from_numpy(g["gap_out"])).abs().max().item() < 2e-2        # sigmoid outputs in (0, 1); bf16 convs
    with pytest.raises(NotImplementedError):
        head.train()(x.cuda())


def test_algebraic_triangulation_net_forward_and_backward():
    """AlgebraicTriangulationNet mirror (lib/models/triangulation.py:217-274): shapes, frozen layers, values against the
    oracle chain (backbone oracle -> soft-argmax -> scaling -> DLT oracle), and a 3-D loss reaching stage4 / last_layer."""
    from oracle import decode_oracle, fixtures, hrnet_oracle, triangulation_oracle as T
    from hrnet_b200.models.triangulation import AlgebraicTriangulationNet
    b, v, H, W = 2, 4, 128, 128
    net = AlgebraicTriangulationNet.from_widths(32, image_size=(H, W))
    frozen = [n for n, p in net.backbone.named_parameters() if not p.requires_grad]
    free = [n for n, p in net.backbone.named_parameters() if p.requires_grad]
    assert all(n.startswith(("stage4.", "last_layer.")) for n in free) and any(n.startswith("stage3.") for n in frozen)
    imgs = fixtures.images(b * v, H, W).view(b, v, 3, H, W)
    P = fixtures.cameras(b, v)
    net.eval()
    torch.manual_seed(11)
    with torch.no_grad():
        k3, k2, hm, conf = net(imgs.cuda(), P.cuda())
    assert k3.shape == (b, 21, 3) and k2.shape == (b, v, 21, 2) and hm.shape == (b, v, 21, H // 4, W // 4) and conf is None
    sd = {k: t.detach().cpu() for k, t in net.backbone.state_dict().items()}
    arch = hrnet_oracle.Arch((32, 64, 128, 256))
    o_heat = hrnet_oracle.forward(sd, imgs.view(-1, 3, H, W), arch, "softmax")[0]
    o_k2 = decode_oracle.spatial_expectation2d(o_heat.numpy()).reshape(b, v, 21, 2) * np.array([640 / (W // 4), 480 / (W // 4)], np.float32)
    assert np.abs(k2.cpu().numpy() - o_k2).max() < 0.05 * 640 / (W // 4)            # 0.05 heat-map px in image units
    bk0 = T.start_vectors(b, 21, 11)
    o_k3 = T.triangulate_joints(k2.cpu().numpy(), P.numpy(), bk0)
    assert np.abs(k3.cpu().numpy() - o_k3).max() < 1e-3 * np.abs(o_k3).max()
    # training: 3-D loss -> stage4 / last_layer gradients through DLT adjoint + soft-argmax + backbone backward
    net.train()
    k3, *_ = net(imgs.cuda(), P.cuda())
    k3.pow(2).mean().backward()
    gl = net.backbone.last_layer[3].weight.grad
    assert gl is not None and torch.isfinite(gl).all() and float(gl.abs().max()) > 0
    assert net.backbone.stage4[0].branches[0][0].conv1.weight.grad is not None
    assert net.backbone.conv1.weight.grad is None


@pytest.mark.parametrize("size,B", [(16, 3), (64, 2)])
def test_cross_view_aggregation_matches_oracle(size, B):
    """row f2: Aggregation (12 ChannelWiseFC GEMMs + weighted fusion) on the tcgen05 GEMM path against the fp32 oracle
    (pinned to the reference's Aggregation by make_golden); 64 x 64 heat maps = the 4096 x 4096 layers of the MHP configs"""
    from oracle import glue_oracle as G
    from hrnet_b200.models.multiview_pose_hrnet import Aggregation
    torch.manual_seed(3)
    ag = Aggregation({"MODEL": {"HEATMAP_SIZE": [size, size]}}).cuda().eval()
    gen_ = torch.Generator().manual_seed(4)
    views = [torch.softmax(torch.randn(B, 21, size * size, generator=gen_) * 3, -1).view(B, 21, size, size) for _ in range(4)]
    ref = G.aggregation([m.weight.weight.detach().cpu() for m in ag.aggre], views)
    with torch.no_grad():
        out = ag([v.cuda() for v in views])
    for a, b in zip(out, ref):
        assert a.shape == b.shape
        assert (a.cpu() - b).abs().max().item() <= 2e-2 * b.abs().max().item()
    with pytest.raises(NotImplementedError):
        ag.train()([v.cuda() for v in views])
