"""-m gpu: every CUDA kernel behind the C ABI against a PyTorch fp32 reference of the same op (convs,
elementwise) or the numpy oracle + reference-generated golden vectors (decode, losses)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf16(t):
    return t.to(torch.bfloat16).float()


def _conv_case(N, H, W, cin, cout, k, stride, relu, use_res, mb=None, nchw=False, force_gather=False, seed=0):
    from hrnet_b200.ops import ConvLayer, PF8
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(N, cin, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(cout, device="cuda", generator=g) * 0.1
    Ho, Wo = H // stride, W // stride
    res = torch.randn(N, cout, Ho, Wo, device="cuda", generator=g) if use_res else None
    layer = ConvLayer(w, scale, shift, stride=stride, relu=relu, out_nchw=nchw)
    layer.force_gather = force_gather
    xp = PF8.from_nchw(x)
    rp = PF8.from_nchw(res) if use_res else None
    if nchw:
        out = torch.full((N, cout, Ho, Wo), float("nan"), device="cuda")
        layer(xp, out, rp, mb=mb)
        got = out
    else:
        op = PF8(N, cout, Ho, Wo)
        op.buf.fill_(7.0)     # poison: padding must be rewritten to zero by the kernel
        op.buf[:, :op.lead] = 0
        op.buf[:, op.lead + op.P:] = 0
        layer(xp, op, rp, mb=mb)
        got = op.to_nchw()
        assert op.padding_is_zero(), "conv must keep PF8 padding/guards zero"
    ref = F.conv2d(_bf16(x), _bf16(w * scale.view(-1, 1, 1, 1)), None, stride=stride, padding=k // 2)
    ref = ref + shift.view(1, -1, 1, 1)
    if use_res:
        ref = ref + _bf16(res)
    if relu:
        ref = F.relu(ref)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    tol = 2e-2 * max(1.0, ref.abs().max().item()) if not nchw else 2e-3 * max(1.0, ref.abs().max().item())
    return err, tol


CONV_CASES = [
    # N, H, W, cin, cout, k, stride, relu, res, mb, nchw, force_gather
    (2, 16, 16, 64, 64, 1, 1, False, False, 1, False, False),      # plain 1x1 GEMM
    (2, 16, 16, 64, 64, 1, 1, True, True, 2, False, False),
    (2, 64, 64, 32, 32, 3, 1, True, True, 1, False, False),        # dominant HRNet shape
    (2, 64, 64, 32, 32, 3, 1, True, True, 2, False, False),
    (1, 64, 64, 32, 32, 3, 1, True, False, 4, False, False),
    (3, 32, 32, 64, 64, 3, 1, True, True, None, False, False),
    (2, 16, 16, 128, 128, 3, 1, True, True, 2, False, False),      # 2 K chunks
    (4, 8, 8, 256, 256, 3, 1, True, True, 1, False, False),        # 4 K chunks, ring wrap
    (2, 64, 64, 256, 32, 3, 1, True, False, None, False, False),   # transition1.0
    (2, 64, 64, 64, 256, 1, 1, True, True, None, False, False),    # bottleneck conv3
    (2, 32, 32, 32, 64, 3, 2, False, False, 1, False, False),      # stride-2 fuse conv (gather)
    (2, 64, 64, 64, 64, 3, 2, True, False, 2, False, False),       # stem conv2 (gather)
    (2, 16, 16, 256, 64, 3, 2, True, False, None, False, False),
    (2, 16, 16, 64, 64, 1, 1, False, False, 1, False, True),       # 1x1 through the gather producer
    (2, 16, 16, 32, 32, 3, 1, True, True, 1, False, True),         # 3x3 s1 through the gather producer
    (2, 32, 32, 480, 480, 1, 1, True, False, None, False, False),  # head conv, 2 N tiles of 240, KC=6
    (2, 32, 32, 480, 21, 1, 1, False, False, None, True, False),   # final conv, fp32 NCHW output
    (2, 16, 16, 96, 21, 3, 1, False, False, None, True, False),    # FINAL_CONV_KERNEL = 3
    (2, 24, 16, 48, 48, 3, 1, True, True, None, False, False),     # W48 widths, non-square
    (2, 12, 8, 96, 192, 3, 2, False, False, None, False, False),
    (1, 24, 16, 720, 720, 1, 1, True, False, None, False, False),  # W48 head, 3 N tiles of 240
    (2, 16, 16, 384, 384, 3, 1, True, True, None, False, False),   # BN=192 x 2
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "N%d_%dx%d_c%d-%d_k%d_s%d_r%d_res%d_mb%s_nchw%d_g%d" % c)
def test_conv_matches_torch(case):
    err, tol = _conv_case(*case)
    assert err <= tol, (err, tol)


def test_stem_conv1():
    from hrnet_b200.ops import PF8, stem_conv1
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(2, 3, 64, 96, device="cuda", generator=g)
    w = torch.randn(64, 3, 3, 3, device="cuda", generator=g) * 0.2
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    out = PF8(2, 64, 32, 48)
    stem_conv1(x, w.reshape(64, 27).contiguous(), b, out)
    ref = F.relu(F.conv2d(x, w, b, stride=2, padding=1))
    assert (out.to_nchw() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert out.padding_is_zero()


def test_layout_roundtrip():
    from hrnet_b200.ops import PF8
    x = torch.randn(3, 40, 12, 20, device="cuda")
    t = PF8.from_nchw(x)
    assert torch.equal(t.to_nchw(), _bf16(x))
    assert torch.equal(t.torch_interior(), _bf16(x))
    assert t.padding_is_zero()


def test_fuse_sum_matches_torch():
    from hrnet_b200.ops import PF8, fuse_sum
    N, C, H, W = 2, 32, 32, 16
    a = torch.randn(N, C, H, W, device="cuda")
    b = torch.randn(N, C, H // 2, W // 2, device="cuda")
    c = torch.randn(N, C, H // 4, W // 4, device="cuda")
    d = torch.randn(N, C, H // 8, W // 8, device="cuda")
    out = PF8(N, C, H, W)
    fuse_sum([PF8.from_nchw(t) for t in (a, b, c, d)], [0, 1, 2, 3], out, relu=True)
    ref = _bf16(a)
    for t, s in ((b, 2), (c, 4), (d, 8)):
        ref = ref + F.interpolate(_bf16(t), scale_factor=s, mode="nearest")
    ref = F.relu(ref)
    assert (out.to_nchw() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    assert out.padding_is_zero()


@pytest.mark.parametrize("align", [True, False])
def test_bilinear_matches_torch(align):
    from hrnet_b200.ops import PF8, bilinear_up
    N, C = 2, 64
    src = torch.randn(N, C, 8, 6, device="cuda")
    cat = PF8(N, 32 + C, 32, 24)
    dst = cat.view_planes(4, C // 8)
    bilinear_up(PF8.from_nchw(src), dst, align)
    ref = F.interpolate(_bf16(src), size=(32, 24), mode="bilinear", align_corners=align)
    got = dst.to_nchw()
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert (cat.view_planes(0, 4).to_nchw() == 0).all()


# ---------------------------------------------------------------------------------------------------
# decode / loss: oracle + golden vectors
# ---------------------------------------------------------------------------------------------------
def _gold(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_decode_against_golden(golden_dir):
    from hrnet_b200.core import inference
    from hrnet_b200.utils import heatmap_decoding
    from hrnet_b200.config import to_cfg
    g = _gold(golden_dir, "decode.npz")
    for tag in ("sq", "rect", "j20"):
        for nm in ("soft", "edge"):
            a = g["%s_%s_in" % (tag, nm)]
            p, mv = inference.get_max_preds(a.copy())
            assert np.array_equal(p, g["%s_%s_maxpreds" % (tag, nm)])          # bit-exact indices
            assert np.array_equal(mv, g["%s_%s_maxvals" % (tag, nm)])
            t = torch.from_numpy(a).cuda()
            assert np.array_equal(heatmap_decoding.get_final_preds(t, False).cpu().numpy(), g["%s_%s_hstride" % (tag, nm)])
            e = heatmap_decoding.get_final_preds(t, True).cpu().numpy()
            assert np.allclose(e, g["%s_%s_expect" % (tag, nm)], rtol=1e-4, atol=1e-3)
            for pp in (0, 1):
                cfg = to_cfg({"TEST": {"POST_PROCESS": bool(pp)}})
                fp, fmv = inference.get_final_preds(cfg, a.copy(), g[tag + "_center"], g[tag + "_scale"])
                assert np.allclose(fp, g["%s_%s_final%d" % (tag, nm, pp)], rtol=1e-5, atol=1e-4)
                assert np.array_equal(fmv, g["%s_%s_maxvals" % (tag, nm)])


def test_softmax_softargmax_against_golden_and_oracle(golden_dir):
    from hrnet_b200 import _lib
    from oracle import decode_oracle
    g = _gold(golden_dir, "decode.npz")
    for tag in ("sq", "rect", "j20"):
        logits = torch.from_numpy(g[tag + "_logits"]).cuda()
        B, J, h, w = logits.shape
        heat = torch.empty_like(logits)
        coords = torch.empty(B, J, 2, device="cuda")
        temp = torch.tensor([1.7], device="cuda")
        _lib.check(_lib.lib().hrnb_softmax_softargmax(logits.data_ptr(), temp.data_ptr(), B * J, h, w, heat.data_ptr(),
                                                      coords.data_ptr(), _lib.stream_ptr()))
        assert np.allclose(heat.cpu().numpy(), g[tag + "_softmax17"], rtol=1e-4, atol=1e-9)
        exp = decode_oracle.spatial_expectation2d(g[tag + "_softmax17"])
        assert np.abs(coords.cpu().numpy() - exp).max() < 1e-3     # << the 0.05 px budget


def test_softargmax_backward_matches_autograd():
    from hrnet_b200.utils.heatmap_decoding import get_final_preds
    hm = torch.rand(2, 21, 16, 12, device="cuda", requires_grad=True)
    out = get_final_preds(hm, True)
    wgt = torch.randn_like(out)
    (out * wgt).sum().backward()
    xs = torch.arange(12, device="cuda", dtype=torch.float32).view(1, 1, 1, 12)
    ys = torch.arange(16, device="cuda", dtype=torch.float32).view(1, 1, 16, 1)
    ref = wgt[..., 0, None, None] * xs + wgt[..., 1, None, None] * ys
    assert torch.allclose(hm.grad, ref.expand_as(hm), atol=1e-5)


def test_softmax_backward_matches_autograd():
    from hrnet_b200 import _lib
    B, J, h, w = 2, 5, 16, 12
    logits = torch.randn(B, J, h, w, device="cuda", requires_grad=True)
    temp = torch.tensor(1.3, device="cuda", requires_grad=True)
    heat_ref = torch.softmax(logits.reshape(B, J, -1) * temp, 2).reshape(B, J, h, w)
    xs = torch.arange(w, device="cuda", dtype=torch.float32).view(1, 1, 1, w)
    ys = torch.arange(h, device="cuda", dtype=torch.float32).view(1, 1, h, 1)
    coords_ref = torch.stack(((heat_ref * xs).sum((2, 3)), (heat_ref * ys).sum((2, 3))), -1)
    dheat = torch.randn_like(heat_ref)
    dcoords = torch.randn_like(coords_ref)
    ((heat_ref * dheat).sum() + (coords_ref * dcoords).sum()).backward()
    d_logits = torch.empty_like(logits)
    d_temp = torch.zeros(1, device="cuda")
    t1 = temp.detach().reshape(1).clone()
    lg = logits.detach()
    _lib.check(_lib.lib().hrnb_softmax_softargmax_bwd(lg.data_ptr(), t1.data_ptr(), heat_ref.detach().contiguous().data_ptr(),
                                                      dheat.data_ptr(), dcoords.data_ptr(), B * J, h, w,
                                                      d_logits.data_ptr(), d_temp.data_ptr(), _lib.stream_ptr()))
    assert torch.allclose(d_logits, logits.grad, rtol=1e-3, atol=1e-6)
    assert torch.allclose(d_temp[0], temp.grad, rtol=1e-3, atol=1e-4)


def test_losses_against_golden(golden_dir):
    from hrnet_b200.core.loss import HeatmapLoss, JointsMSELoss
    g = _gold(golden_dir, "loss.npz")
    for tag in ("a", "b"):
        gt = torch.from_numpy(g[tag + "_gt"]).cuda()
        for mode in ("l2", "l1"):
            pred = torch.from_numpy(g[tag + "_pred"]).cuda().requires_grad_(True)
            l = HeatmapLoss(mode)(pred, gt)
            l.backward()
            assert np.allclose(l.item(), g["%s_hm_%s" % (tag, mode)], rtol=1e-5)
            assert np.allclose(pred.grad.cpu().numpy(), g["%s_hm_%s_grad" % (tag, mode)], rtol=1e-5, atol=1e-8)
        xy = torch.from_numpy(g[tag + "_xy"]).cuda()
        vis = torch.from_numpy(g[tag + "_vis"]).cuda()
        for vtag, v in (("vis", vis), ("novis", None), ("zerovis", torch.zeros_like(vis))):
            pp = torch.from_numpy(g[tag + "_pp"]).cuda().requires_grad_(True)
            l = JointsMSELoss()(pp, xy, v)
            l.backward()
            assert np.allclose(l.item(), g["%s_p2d_%s" % (tag, vtag)], rtol=1e-5)
            assert np.allclose(pp.grad.cpu().numpy(), g["%s_p2d_%s_grad" % (tag, vtag)], rtol=1e-4, atol=1e-7)


def test_decode_full_size_properties():
    """BASELINE-size maps (B=64, 21 joints, 64x64): size-independent properties."""
    from hrnet_b200.core.inference import get_max_preds_cuda
    from hrnet_b200.utils.heatmap_decoding import get_final_preds
    B, J, h, w = 64, 21, 64, 64
    hm = torch.rand(B, J, h, w, device="cuda")
    yy = torch.randint(0, h, (B, J), device="cuda")
    xx = torch.randint(0, w, (B, J), device="cuda")
    hm[torch.arange(B)[:, None], torch.arange(J)[None, :], yy, xx] = 2.0     # planted unique maxima
    p, mv = get_max_preds_cuda(hm)
    assert torch.equal(p[..., 0].long(), xx) and torch.equal(p[..., 1].long(), yy)
    assert (mv == 2.0).all()
    idx = hm.reshape(B, J, -1).argmax(2)
    assert torch.equal(get_final_preds(hm, False), torch.stack((idx % h, idx // h), 2).float())
    # expectation is linear: E[a*p + b*q] = a*E[p] + b*E[q]
    q = torch.rand_like(hm)
    lhs = get_final_preds(0.3 * hm + 0.7 * q, True)
    rhs = 0.3 * get_final_preds(hm, True) + 0.7 * get_final_preds(q, True)
    assert torch.allclose(lhs, rhs, rtol=1e-4, atol=1e-1)


@pytest.mark.parametrize("N,H,W,cin,cout,relu", [(2, 32, 32, 32, 64, False), (2, 64, 64, 64, 64, True), (3, 16, 24, 128, 256, False),
                                                 (1, 16, 16, 256, 64, True), (2, 24, 16, 48, 96, True)])
def test_stride2_conv_over_phase_split_input(N, H, W, cin, cout, relu):
    """3x3 stride-2 conv on the flat-shift path: phase_split -> HRNB_CONV_IN_PHASES conv, against torch."""
    from hrnet_b200.ops import ConvLayer, PF8, PhasePF8, phase_split
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(N, cin, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(cout, device="cuda", generator=g) * 0.1
    xp = PF8.from_nchw(x)
    ph = PhasePF8(N, cin, H, W)
    phase_split(xp, ph)
    assert torch.equal(ph.to_nchw(), _bf16(x)) and ph.padding_is_zero()
    layer = ConvLayer(w, scale, shift, stride=2, relu=relu)
    out = PF8(N, cout, H // 2, W // 2)
    layer(ph, out)
    ref = F.conv2d(_bf16(x), _bf16(w * scale.view(-1, 1, 1, 1)), None, stride=2, padding=1) + shift.view(1, -1, 1, 1)
    if relu:
        ref = F.relu(ref)
    assert (out.to_nchw() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
    assert out.padding_is_zero()


def test_conv_writes_phase_split_output_for_a_following_stride2_conv():
    from hrnet_b200.ops import ConvLayer, PF8, PhasePF8
    g = torch.Generator(device="cuda").manual_seed(12)
    N, C, H, W = 2, 32, 32, 32
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    w1 = torch.randn(C, C, 3, 3, device="cuda", generator=g) / (C * 9) ** 0.5
    w2 = torch.randn(64, C, 3, 3, device="cuda", generator=g) / (C * 9) ** 0.5
    ph = PhasePF8(N, C, H, W)
    ConvLayer(w1, relu=True)(PF8.from_nchw(x), ph)                    # stride-1 conv, OUT_PHASES epilogue
    y_ref = F.relu(F.conv2d(_bf16(x), _bf16(w1), None, padding=1))
    assert (ph.to_nchw() - y_ref).abs().max().item() <= 2e-2 * y_ref.abs().max().item()
    assert ph.padding_is_zero()
    out = PF8(N, 64, H // 2, W // 2)
    ConvLayer(w2, stride=2)(ph, out)
    z_ref = F.conv2d(ph.to_nchw(), _bf16(w2), None, stride=2, padding=1)
    assert (out.to_nchw() - z_ref).abs().max().item() <= 2e-2 * z_ref.abs().max().item()


def test_stem_im2col_then_1x1_equals_conv1():
    from hrnet_b200.ops import ConvLayer, PF8, stem_im2col
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(2, 3, 64, 96, device="cuda", generator=g)
    w = torch.randn(64, 3, 3, 3, device="cuda", generator=g) * 0.2
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    cols = PF8(2, 32, 32, 48)
    stem_im2col(x, cols)
    assert cols.padding_is_zero()
    w32 = torch.zeros(64, 32, 1, 1, device="cuda")
    w32[:, :27, 0, 0] = w.reshape(64, 27)
    out = PF8(2, 64, 32, 48)
    ConvLayer(w32, None, b, relu=True)(cols, out)
    ref = F.relu(F.conv2d(_bf16(x), _bf16(w), b, stride=2, padding=1))
    assert (out.to_nchw() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("N,H,W,cin,cout,k,stride,shifts", [
    (2, 32, 32, 32, 64, 3, 2, (1, 2)),          # stage-4 output 1: host = stride-2 conv from branch 0, + up-sampled branches 2, 3
    (2, 16, 16, 64, 128, 3, 2, (0, 1)),         # output 2: + another chain's output (same grid) + up-sampled branch 3
    (3, 8, 8, 128, 256, 3, 2, (0, 0)),          # output 3: + two other chains
    (2, 16, 24, 64, 64, 3, 1, (0, 1, 3)),       # three sources on a stride-1 host, non-square
])
def test_conv_with_fuse_sources_in_the_epilogue(N, H, W, cin, cout, k, stride, shifts):
    """include/hrnb.h nfuse: out = ReLU(conv(x) + bias + res + sum_f nearest_up(src_f)) in one launch - the fuse-layer sum of
    HighResolutionModule.forward (lib/models/pose_hrnet.py:257-266) inside the epilogue of the output's own stride-2 conv"""
    from hrnet_b200.ops import ConvLayer, PF8, PhasePF8, phase_split
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(N, cin, H * stride, W * stride, device="cuda", generator=g)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(cout, device="cuda", generator=g) * 0.1
    res = torch.randn(N, cout, H, W, device="cuda", generator=g)
    srcs = [torch.randn(N, cout, H >> s, W >> s, device="cuda", generator=g) for s in shifts]
    layer = ConvLayer(w, scale, shift, stride=stride, relu=False)
    xp = PF8.from_nchw(x)
    if stride == 2:
        ph = PhasePF8(N, cin, H * 2, W * 2)
        phase_split(xp, ph)
        xp = ph
    out = PF8(N, cout, H, W)
    out.buf.fill_(7.0); out.buf[:, :out.lead] = 0; out.buf[:, out.lead + out.P:] = 0
    layer(xp, out, PF8.from_nchw(res), fuse=[(PF8.from_nchw(t), s) for t, s in zip(srcs, shifts)], relu=True)
    ref = F.conv2d(_bf16(x), _bf16(w * scale.view(-1, 1, 1, 1)), None, stride=stride, padding=k // 2) + shift.view(1, -1, 1, 1)
    ref = ref + _bf16(res)
    for t, s in zip(srcs, shifts):
        ref = ref + F.interpolate(_bf16(t), scale_factor=2 ** s, mode="nearest") if s else ref + _bf16(t)
    ref = F.relu(ref)
    torch.cuda.synchronize()
    assert out.padding_is_zero()
    assert (out.to_nchw() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("N,H,W,cin,cout,up,shifts", [(2, 64, 64, 64, 32, 1, (2, 3)), (2, 32, 32, 128, 64, 1, ()), (1, 32, 48, 96, 48, 2, (1,))])
def test_1x1_conv_on_upsampled_input_hosts_the_fuse_sum(N, H, W, cin, cout, up, shifts):
    """include/hrnb.h in_up_shift: the fuse output of the highest-resolution branch - the 1x1 conv from the next branch
    evaluated on the fine grid (input read through nearest up-sampling), identity as residual, other up paths as fuse sources"""
    from hrnet_b200.ops import ConvLayer, PF8
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(N, cin, H >> up, W >> up, device="cuda", generator=g)
    w = torch.randn(cout, cin, 1, 1, device="cuda", generator=g) / cin ** 0.5
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(cout, device="cuda", generator=g) * 0.1
    res = torch.randn(N, cout, H, W, device="cuda", generator=g)
    srcs = [torch.randn(N, cout, H >> s, W >> s, device="cuda", generator=g) for s in shifts]
    layer = ConvLayer(w, scale, shift, relu=False)
    out = PF8(N, cout, H, W)
    out.buf.fill_(7.0); out.buf[:, :out.lead] = 0; out.buf[:, out.lead + out.P:] = 0
    layer(PF8.from_nchw(x), out, PF8.from_nchw(res), fuse=[(PF8.from_nchw(t), s) for t, s in zip(srcs, shifts)], relu=True,
          up_shift=up)
    z = F.conv2d(_bf16(x), _bf16(w * scale.view(-1, 1, 1, 1))) + shift.view(1, -1, 1, 1)
    ref = F.interpolate(z, scale_factor=2 ** up, mode="nearest") + _bf16(res)
    for t, s in zip(srcs, shifts):
        ref = ref + F.interpolate(_bf16(t), scale_factor=2 ** s, mode="nearest")
    ref = F.relu(ref)
    torch.cuda.synchronize()
    assert out.padding_is_zero()
    assert (out.to_nchw() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("N,H,W,C,shifts", [(2, 64, 64, 32, (1, 2, 3)), (3, 32, 48, 64, (1,)), (2, 16, 16, 128, (1, 2))])
def test_last_branch_conv_hosts_fuse_output_0(N, H, W, C, shifts):
    """HRNB_CONV_FUSE_AFTER_RELU + out2: the last 3x3 conv of branch 0 writes its own output x0 = ReLU(conv + bias + res) as the
    phase-split copy the stride-2 chains read, and ReLU(x0 + sum_j nearest_up(z_j)) - fuse output 0 of
    HighResolutionModule.forward (lib/models/pose_hrnet.py:257-266) - as the primary output, in one launch"""
    from hrnet_b200.ops import ConvLayer, PF8, PhasePF8
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / (C * 9) ** 0.5
    scale = torch.rand(C, device="cuda", generator=g) + 0.5
    shift = torch.randn(C, device="cuda", generator=g) * 0.1
    res = torch.randn(N, C, H, W, device="cuda", generator=g)
    srcs = [torch.randn(N, C, H >> s, W >> s, device="cuda", generator=g) for s in shifts]
    out, ph = PF8(N, C, H, W), PhasePF8(N, C, H, W)
    out.buf.fill_(7.0); out.buf[:, :out.lead] = 0; out.buf[:, out.lead + out.P:] = 0
    ConvLayer(w, scale, shift, relu=True)(PF8.from_nchw(x), out, PF8.from_nchw(res), out2=ph,
                                          fuse=[(PF8.from_nchw(t), s) for t, s in zip(srcs, shifts)], fuse_after=True)
    x0 = F.relu(F.conv2d(_bf16(x), _bf16(w * scale.view(-1, 1, 1, 1)), None, padding=1) + shift.view(1, -1, 1, 1) + _bf16(res))
    ref = x0
    for t, s in zip(srcs, shifts):
        ref = ref + F.interpolate(_bf16(t), scale_factor=2 ** s, mode="nearest")
    ref = F.relu(ref)
    torch.cuda.synchronize()
    assert out.padding_is_zero() and ph.padding_is_zero()
    assert (ph.to_nchw() - x0).abs().max().item() <= 2e-2 * max(1.0, x0.abs().max().item())
    assert (out.to_nchw() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


def test_lean_epilogue_phase_copy_equals_plain_output():
    """out2 on the lean epilogue: the phase-split copy holds exactly the bf16 values of the plain PF8 output"""
    from hrnet_b200.ops import ConvLayer, PF8, PhasePF8
    g = torch.Generator(device="cuda").manual_seed(9)
    for N, C, H, W in ((2, 32, 64, 64), (3, 64, 32, 16), (1, 256, 16, 16)):
        x = torch.randn(N, C, H, W, device="cuda", generator=g)
        w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / (C * 9) ** 0.5
        res = torch.randn(N, C, H, W, device="cuda", generator=g)
        out, ph = PF8(N, C, H, W), PhasePF8(N, C, H, W)
        ConvLayer(w, relu=True)(PF8.from_nchw(x), out, PF8.from_nchw(res), out2=ph)
        torch.cuda.synchronize()
        assert torch.equal(ph.to_nchw(), out.to_nchw()) and ph.padding_is_zero() and out.padding_is_zero()


@pytest.mark.parametrize("N,C,H,W,use_res,nfuse", [(160, 32, 64, 64, True, 0), (160, 32, 64, 64, False, 0), (128, 48, 96, 72, True, 0),
                                                   (160, 32, 64, 64, True, 2)])
def test_two_ctas_per_sm_variant_equals_the_single_cta_launch(N, C, H, W, use_res, nfuse):
    """conv_tc.cu `twin` (opt-in, hrnb_debug_set(9, n)): thin resident-weight layers with >= n tiles per CTA slot run the
    8-epilogue-warp instantiation with two CTAs per SM - bit-identical to the 16-warp launch and right vs torch"""
    from hrnet_b200 import _lib
    from hrnet_b200.ops import ConvLayer, PF8
    g = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / (C * 9) ** 0.5
    shift = torch.randn(C, device="cuda", generator=g) * 0.1
    res = PF8.from_nchw(torch.randn(N, C, H, W, device="cuda", generator=g)) if use_res else None
    srcs = [torch.randn(N, C, H >> s, W >> s, device="cuda", generator=g) for s in range(1, nfuse + 1)]
    fuse = [(PF8.from_nchw(t), s + 1) for s, t in enumerate(srcs)] or None
    layer, xp = ConvLayer(w, None, shift, relu=True), PF8.from_nchw(x)
    outs = []
    for off in (0, 1):
        _lib.check(_lib.lib().hrnb_debug_set(9, 0 if off else 2))
        try:
            o = PF8(N, C, H, W)
            layer(xp, o, res, fuse=fuse)
            torch.cuda.synchronize()
            outs.append(o)
        finally:
            _lib.lib().hrnb_debug_set(9, 0)
    assert torch.equal(outs[0].buf, outs[1].buf) and outs[0].padding_is_zero()
    ref = F.conv2d(_bf16(x), _bf16(w), None, padding=1) + shift.view(1, -1, 1, 1)
    if use_res:
        ref = ref + res.to_nchw()
    for s, t in enumerate(srcs):
        ref = ref + F.interpolate(_bf16(t), scale_factor=2 ** (s + 1), mode="nearest")
    ref = F.relu(ref)
    assert (outs[0].to_nchw() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


def test_fuse_in_epilogue_plan_equals_stand_alone_sum_plan():
    """whole network: the inference plans with the fuse sums inside conv epilogues (output 0 hosted by branch 0's last conv - up-path
    terms pre-summed on the low-resolution grids or as separate sources - or by the gathered 1x1 conv) against the plan with
    fuse_sum kernels"""
    import os
    from oracle import fixtures
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet_softmax
    outs = []
    for flag, host0, tree in (("1", "conv2", "1"), ("1", "conv2", "0"), ("1", "gather", "1"), ("0", "conv2", "1")):
        os.environ["HRNB_FUSE_EPILOGUE"], os.environ["HRNB_FUSE_HOST0"], os.environ["HRNB_FUSE_TREE"] = flag, host0, tree
        try:
            torch.manual_seed(0)
            m = pose_hrnet_softmax.get_pose_net(make_cfg(32), is_train=False)
            sd = m.state_dict(); fixtures.perturb_state_dict(sd); m.load_state_dict(sd)
            m = m.cuda().eval()
            h, f, _ = m(fixtures.images(2, 128, 128).cuda())
            outs.append((h.clone(), f.clone(), m.engine().plan(2, 128, 128).launches(False)))
        finally:
            os.environ.pop("HRNB_FUSE_EPILOGUE", None)
            os.environ.pop("HRNB_FUSE_HOST0", None)
            os.environ.pop("HRNB_FUSE_TREE", None)
    assert all(o[2] < outs[3][2] for o in outs[:3])                 # the fuse_sum launches are gone
    for a in outs[:3]:
        assert (a[1] - outs[3][1]).abs().max().item() <= 4e-2 * outs[3][1].abs().max().item()
        assert (a[0] - outs[3][0]).abs().max().item() <= 2e-2 * outs[3][0].abs().max().item()
