"""-m gpu: the TRAINING kernels behind the C ABI (wgrad on tcgen05, data-gradient convs through the conv kernel's
transposed / custom-tap packing, batch-stat BN forward + backward, fuse / bilinear / phase backward, fused Adam)
against PyTorch fp32 autograd of the same op on the same bf16-rounded inputs."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf16(t):
    return t.to(torch.bfloat16).float()


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale


def _relerr(got, ref):
    return (got - ref).abs().max().item() / max(1e-12, ref.abs().max().item())


WGRAD_CASES = [
    # N, H, W, cin, cout, k, stride, overrides
    (2, 16, 16, 32, 32, 1, 1, {}),
    (2, 16, 16, 32, 32, 3, 1, {}),
    (2, 64, 64, 32, 32, 3, 1, {}),
    (3, 32, 32, 64, 64, 3, 1, {}),
    (2, 16, 16, 128, 128, 3, 1, {}),
    (4, 8, 8, 256, 256, 3, 1, {}),
    (2, 64, 64, 256, 32, 3, 1, {}),
    (2, 32, 32, 64, 256, 1, 1, {}),
    (2, 32, 32, 480, 480, 1, 1, {}),
    (2, 32, 32, 480, 21, 1, 1, {}),
    (2, 32, 32, 32, 64, 1, 1, {}),            # stem conv1 as 1x1 over the im2col slab
    (2, 24, 16, 48, 96, 3, 1, {}),            # W48 widths, non-square
    (2, 16, 16, 32, 32, 3, 1, dict(KP=64, ksplit=3, TG=2)),
    (2, 16, 16, 64, 64, 3, 1, dict(NT=16, KP=32)),
    (2, 32, 32, 32, 64, 3, 2, {}),
    (2, 64, 64, 64, 64, 3, 2, {}),
    (3, 16, 24, 128, 256, 3, 2, {}),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: "N%d_%dx%d_c%d-%d_k%d_s%d" % c[:7] + ("_ovr" if c[7] else ""))
def test_wgrad_matches_autograd(case):
    from hrnet_b200 import tops
    from hrnet_b200.ops import PF8, PhasePF8, phase_split
    N, H, W, cin, cout, k, stride, ovr = case
    x = _bf16(_rand(N, cin, H, W, seed=1))
    Ho, Wo = H // stride, W // stride
    dy = _bf16(_rand(N, cout, Ho, Wo, seed=2))
    w = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)
    F.conv2d(x, w, None, stride=stride, padding=k // 2).backward(dy)
    ref = w.grad.permute(2, 3, 1, 0).reshape(k * k, cin, cout)          # [tap][cin][cout]
    dyp = PF8.from_nchw(dy)
    xp = PF8.from_nchw(x)
    dw = torch.zeros(k * k, cin, cout, device="cuda")
    if stride == 2:
        ph = PhasePF8(N, cin, H, W)
        phase_split(xp, ph)
        tops.wgrad_conv(dyp, ph, dw, k, 2)
    elif ovr:
        tops.wgrad(dyp, xp, dw, cin, cout, tops.fwd_taps_s1(k, dyp.Wp), **ovr)
    else:
        tops.wgrad_conv(dyp, xp, dw, k, 1)
    torch.cuda.synchronize()
    assert _relerr(dw, ref) < 2e-3, _relerr(dw, ref)
    # accumulation: a second launch doubles the result
    if stride == 1 and not ovr:
        tops.wgrad_conv(dyp, xp, dw, k, 1)
        assert _relerr(dw, 2 * ref) < 2e-3


DGRAD_CASES = [
    (2, 16, 16, 32, 32, 1, False), (2, 64, 64, 32, 32, 3, True), (3, 32, 32, 64, 64, 3, False), (2, 16, 16, 128, 128, 3, True),
    (4, 8, 8, 256, 256, 3, False), (2, 64, 64, 256, 32, 3, False), (2, 32, 32, 64, 256, 1, True), (2, 32, 32, 480, 21, 1, False),
    (2, 24, 16, 48, 96, 3, False),
]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=lambda c: "N%d_%dx%d_c%d-%d_k%d_acc%d" % c)
def test_dgrad_stride1_matches_autograd(case):
    """dX = conv(dY, W^T with mirrored taps) through hrnb_conv; `acc` adds into an existing gradient (res = out)."""
    from hrnet_b200 import tops
    from hrnet_b200.ops import ConvLayer, PF8
    N, H, W, cin, cout, k, acc = case
    w = _bf16(_rand(cout, cin, k, k, seed=3, scale=1.0 / (cin * k * k) ** 0.5)).contiguous()
    dy = _bf16(_rand(N, cout, H, W, seed=4))
    x = torch.zeros(N, cin, H, W, device="cuda", requires_grad=True)
    F.conv2d(x, w, None, padding=k // 2).backward(dy)
    ref = x.grad
    cpad = (cout + 15) // 16 * 16
    tap_ids, _ = tops.dgrad_taps_s1(k, W + 1)
    layer = ConvLayer(w, transpose=True, tap_ids=tap_ids, cin_pad=cpad)
    dyp = PF8(N, cpad, H, W)
    from hrnet_b200 import _lib
    _lib.check(_lib.lib().hrnb_nchw_f32_to_pf8(dy.data_ptr(), N, cout, H, W, dyp.ptr, dyp.ps, _lib.stream_ptr()))
    dx = PF8(N, cin, H, W)
    if acc:
        prev = _bf16(_rand(N, cin, H, W, seed=5))
        dx = PF8.from_nchw(prev)
        layer(dyp, dx, res=dx)
        ref = ref + prev
    else:
        layer(dyp, dx)
    torch.cuda.synchronize()
    assert _relerr(dx.to_nchw(), ref) < 1.5e-2
    assert dx.padding_is_zero()


@pytest.mark.parametrize("N,H,W,cin,cout", [(2, 32, 32, 32, 64), (2, 64, 64, 64, 64), (3, 16, 24, 128, 256), (1, 16, 16, 256, 64)])
def test_dgrad_stride2_matches_autograd(N, H, W, cin, cout):
    from hrnet_b200 import tops
    from hrnet_b200.ops import ConvLayer, PF8, PhasePF8
    w = _bf16(_rand(cout, cin, 3, 3, seed=6, scale=1.0 / (cin * 9) ** 0.5)).contiguous()
    dy = _bf16(_rand(N, cout, H // 2, W // 2, seed=7))
    x = torch.zeros(N, cin, H, W, device="cuda", requires_grad=True)
    F.conv2d(x, w, None, stride=2, padding=1).backward(dy)
    dyp = PF8.from_nchw(dy)
    dph = PhasePF8(N, cin, H, W)
    half = dph.half
    for ph, (tap_ids, taps) in tops.dgrad_taps_s2(dyp.Wp).items():
        layer = ConvLayer(w, transpose=True, tap_ids=tap_ids, custom_taps=taps)
        out = PF8(N, cin, H // 2, W // 2, buf=dph.buf[ph])
        layer(dyp, out)
    assert dph.padding_is_zero()
    assert _relerr(dph.to_nchw(), x.grad) < 1.5e-2
    prev = _bf16(_rand(N, cin, H, W, seed=8))
    dx = PF8.from_nchw(prev)
    tops.phase_merge(dph, dx, mode=2)
    assert _relerr(dx.to_nchw(), _bf16(dph.to_nchw() + prev)) < 1e-2
    dx2 = PF8(N, cin, H, W)
    dx2.buf.fill_(3.0)
    tops.phase_merge(dph, dx2, mode=1)
    assert torch.equal(dx2.to_nchw(), dph.to_nchw())


def test_bn_apply_writes_the_phase_split_copy():
    """hrnb_bn_params.out2: the normalisation kernel also writes the unit output in the phase-split form the stride-2 convs
    read - bit-identical to phase_split of the plain output, padding untouched"""
    from hrnet_b200 import tops
    from hrnet_b200.ops import PF8, PhasePF8, phase_split
    for N, C_, H, W, use_res in ((2, 32, 16, 16, True), (3, 64, 8, 12, False), (64, 32, 64, 64, True)):
        c = PF8.from_nchw(_bf16(_rand(N, C_, H, W, seed=30) * 1.5 + 0.3))
        res = PF8.from_nchw(_bf16(_rand(N, C_, H, W, seed=31))) if use_res else None
        gamma, beta = torch.rand(C_, device="cuda") + 0.5, torch.randn(C_, device="cuda") * 0.1
        sums = torch.zeros(C_, 2, device="cuda")
        tops.bn_stats(c, sums)
        y, y2, ph, ref = PF8(N, C_, H, W), PF8(N, C_, H, W), PhasePF8(N, C_, H, W), PhasePF8(N, C_, H, W)
        tops.bn_apply(c, sums, gamma, beta, y, res=res, relu=True)
        tops.bn_apply(c, sums, gamma, beta, y2, res=res, relu=True, out2=ph)
        phase_split(y, ref)
        torch.cuda.synchronize()
        assert torch.equal(y.buf, y2.buf) and torch.equal(ph.buf, ref.buf) and ph.padding_is_zero()


@pytest.mark.parametrize("N,H,W,cin,cout,accumulate", [(2, 32, 32, 32, 64, False), (2, 64, 64, 64, 64, True), (3, 16, 24, 128, 256, False),
                                                       (1, 16, 16, 256, 64, True), (64, 64, 64, 32, 32, True)])
def test_grouped_dgrad_stride2_equals_per_phase_launches(N, H, W, cin, cout, accumulate):
    """include/hrnb.h ngroup: the four per-phase data-gradient convs of a 3x3 stride-2 conv in ONE launch - equal to
    the four separate launches, writing or accumulating"""
    import ctypes as C
    from hrnet_b200 import _lib, tops
    from hrnet_b200.ops import ConvLayer, PF8, PhasePF8, grouped_conv_params
    w = _bf16(_rand(cout, cin, 3, 3, seed=16, scale=1.0 / (cin * 9) ** 0.5)).contiguous()
    dyp = PF8.from_nchw(_bf16(_rand(N, cout, H // 2, W // 2, seed=17)))
    prev = _bf16(_rand(N, cin, H, W, seed=18))
    ref, got = PhasePF8(N, cin, H, W), PhasePF8(N, cin, H, W)
    if accumulate:
        for d in (ref, got):
            tops_src = PF8.from_nchw(prev)
            from hrnet_b200.ops import phase_split
            phase_split(tops_src, d)
    layers, outs_ref, outs = [], [], []
    for ph, (tap_ids, taps) in sorted(tops.dgrad_taps_s2(dyp.Wp).items()):
        layers.append(ConvLayer(w, transpose=True, tap_ids=tap_ids, custom_taps=taps))
        outs_ref.append(PF8(N, cin, H // 2, W // 2, buf=ref.buf[ph]))
        outs.append(PF8(N, cin, H // 2, W // 2, buf=got.buf[ph]))
    gp = grouped_conv_params(layers, dyp, outs, outs if accumulate else None)
    for ph, layer in enumerate(layers):     # per-phase launches with the grouped launch's tile shape
        layer(dyp, outs_ref[ph], outs_ref[ph] if accumulate else None, mb=gp.MB, bn=gp.BN)
    _lib.check(_lib.lib().hrnb_conv(C.byref(gp), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert got.padding_is_zero()
    assert _relerr(got.to_nchw(), ref.to_nchw()) < 2e-3      # K-chunk size may differ between the two forms: fp32 summation order
    x = torch.zeros(N, cin, H, W, device="cuda", requires_grad=True)
    F.conv2d(x, w, None, stride=2, padding=1).backward(dyp.to_nchw())
    assert _relerr(got.to_nchw(), x.grad + (prev if accumulate else 0)) < 1.5e-2


@pytest.mark.parametrize("N,C,H,W,relu,use_res", [(2, 32, 16, 16, True, True), (4, 64, 8, 12, True, False), (2, 256, 8, 8, False, False),
                                                  (64, 32, 64, 64, True, True), (64, 32, 64, 64, True, False)])
def test_bn_train_forward_backward(N, C, H, W, relu, use_res):
    from hrnet_b200 import tops
    from hrnet_b200.ops import PF8
    c = _bf16(_rand(N, C, H, W, seed=9) * 1.5 + 0.3)
    res = _bf16(_rand(N, C, H, W, seed=10)) if use_res else None
    gamma = (torch.rand(C, device="cuda") + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device="cuda") * 0.2).requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    rm_ref, rv_ref = rm.clone(), rv.clone()
    cr = c.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if use_res else None
    yr = F.batch_norm(cr, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)
    if use_res:
        yr = yr + rr
    if relu:
        yr = F.relu(yr)
    dy = _bf16(_rand(N, C, H, W, seed=11))
    yr.backward(dy)
    cp, rp = PF8.from_nchw(c), (PF8.from_nchw(res) if use_res else None)
    sums = torch.zeros(C, 2, device="cuda")
    tops.bn_stats(cp, sums)
    y = PF8(N, C, H, W)
    y.buf.fill_(5.0)
    y.buf[:, :y.lead] = 0
    y.buf[:, y.lead + y.P:] = 0
    tops.bn_apply(cp, sums, gamma.detach(), beta.detach(), y, res=rp, relu=relu, running_mean=rm, running_var=rv)
    assert y.padding_is_zero()
    assert _relerr(y.to_nchw(), yr.detach()) < 1e-2
    assert torch.allclose(rm, rm_ref, rtol=1e-4, atol=1e-5) and torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-5)
    # backward (dc overwrites dy in place, residual gradient accumulated onto an existing buffer)
    dyp = PF8.from_nchw(dy)
    dsums = torch.zeros(C, 2, device="cuda")
    dgamma, dbeta = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    prev = _bf16(_rand(N, C, H, W, seed=12))
    dres = PF8.from_nchw(prev) if use_res else None
    tops.bn_bwd(dyp, y, cp, sums, gamma.detach(), dsums, dyp, dgamma, dbeta, relu=relu, dres=dres, dres_mode=2)
    assert dyp.padding_is_zero()
    assert _relerr(dyp.to_nchw(), cr.grad) < 2e-2
    assert _relerr(dgamma, gamma.grad) < 1e-2 and _relerr(dbeta, beta.grad) < 1e-2
    if use_res:
        assert _relerr(dres.to_nchw(), rr.grad + prev) < 1e-2


STATS_CASES = [
    # N, H, W, cin, cout, k, stride
    (2, 16, 16, 32, 32, 3, 1),
    (64, 64, 64, 32, 32, 3, 1),      # BASELINE configs[1] branch-0 layer at its full size
    (3, 32, 32, 64, 64, 3, 1),
    (5, 24, 40, 64, 32, 1, 1),
    (2, 32, 32, 32, 16, 1, 1),
    (2, 32, 32, 32, 64, 3, 2),
    (64, 32, 32, 256, 64, 1, 1),
]


@pytest.mark.parametrize("case", STATS_CASES, ids=lambda c: "N%d_%dx%d_c%d-%d_k%d_s%d" % c)
def test_conv_fused_bn_statistics(case):
    """conv launch with stats_sums (BatchNorm batch statistics reduced in the conv epilogue, include/hrnb.h) against
    torch (fp64 sums of the fp32 conv of the same bf16 inputs), against the separate hrnb_bn_stats pass over the bf16
    output, bit-identical when repeated (fixed summation order), and the conv output itself unchanged."""
    from hrnet_b200 import tops
    from hrnet_b200.ops import ConvLayer, PF8, PhasePF8, phase_split
    N, H, W, cin, cout, k, stride = case
    x = _bf16(_rand(N, cin, H, W, seed=21) + 0.2)
    w = _bf16(_rand(cout, cin, k, k, seed=22, scale=1.0 / (cin * k * k) ** 0.5)).contiguous()
    ref = F.conv2d(x, w, None, stride=stride, padding=k // 2).double()
    ref_sums = torch.stack([ref.sum(dim=(0, 2, 3)), (ref * ref).sum(dim=(0, 2, 3))], dim=1)
    xp = PF8.from_nchw(x)
    if stride == 2:
        ph = PhasePF8(N, cin, H, W)
        phase_split(xp, ph)
        xp = ph
    layer = ConvLayer(w, stride=stride)
    Ho, Wo = H // stride, W // stride
    plain, out = PF8(N, cout, Ho, Wo), PF8(N, cout, Ho, Wo)
    layer(xp, plain, bn=cout)
    sums = torch.full((cout, 2), 7.0, device="cuda")
    layer(xp, out, bn=cout, stats=sums)
    torch.cuda.synchronize()
    assert torch.equal(out.buf, plain.buf)
    tol = 2e-3 * ref_sums.abs().max(dim=0).values.float()
    assert ((sums - ref_sums.float()).abs() <= tol).all(), (sums - ref_sums.float()).abs().max(dim=0)
    separate = torch.zeros(cout, 2, device="cuda")
    tops.bn_stats(out, separate)
    assert ((sums - separate).abs() <= tol).all()
    again = torch.zeros(cout, 2, device="cuda")
    layer(xp, out, bn=cout, stats=again)      # a new launch struct = a new workspace
    assert torch.equal(again, sums)
    from hrnet_b200 import _lib
    from hrnet_b200.ops import attach_stats
    p = layer.params(xp, out, bn=cout)
    third = torch.zeros(cout, 2, device="cuda")
    assert attach_stats(p, third)
    for _ in range(3):               # the same struct replayed: the ticket counter must have reset itself
        third.zero_()
        _lib.check(_lib.lib().hrnb_conv(C.byref(p), _lib.stream_ptr()))
    assert torch.equal(third, sums)


def test_conv_fused_bn_statistics_rejects_ineligible_launches():
    from hrnet_b200.ops import ConvLayer, PF8
    x = PF8.from_nchw(_rand(2, 128, 16, 16, seed=23))
    layer = ConvLayer(_rand(128, 128, 3, 3, seed=24, scale=0.03).contiguous())
    with pytest.raises(ValueError):
        layer(x, PF8(2, 128, 16, 16), stats=torch.zeros(128, 2, device="cuda"))


def test_fuse_sum_backward_matches_autograd():
    from hrnet_b200 import tops
    from hrnet_b200.ops import PF8, fuse_sum
    N, C, H, W = 2, 32, 32, 16
    srcs = [_bf16(_rand(N, C, H >> s, W >> s, seed=20 + s)).requires_grad_(True) for s in range(4)]
    tot = srcs[0]
    for s in range(1, 4):
        tot = tot + F.interpolate(srcs[s], scale_factor=2 ** s, mode="nearest")
    yr = F.relu(tot)
    dy = _bf16(_rand(N, C, H, W, seed=30))
    yr.backward(dy)
    y = PF8(N, C, H, W)
    fuse_sum([PF8.from_nchw(t.detach()) for t in srcs], [0, 1, 2, 3], y, relu=True)
    dyp = PF8.from_nchw(dy)
    for s in range(4):
        prev = _bf16(_rand(N, C, H >> s, W >> s, seed=40 + s))
        d = PF8.from_nchw(prev)
        tops.fuse_sum_bwd(dyp, y, d, s, relu=True, mode=2)
        assert _relerr(d.to_nchw(), srcs[s].grad + prev) < 1.5e-2
        d2 = PF8(N, C, H >> s, W >> s)
        tops.fuse_sum_bwd(dyp, y, d2, s, relu=True, mode=1)
        assert _relerr(d2.to_nchw(), srcs[s].grad) < 1.5e-2 and d2.padding_is_zero()
    # all four sources in one launch (hrnb_fuse_sum_bwd_batch), mixed write / accumulate: bit-identical to the single launches
    prevs = [_bf16(_rand(N, C, H >> s, W >> s, seed=40 + s)) for s in range(4)]
    one = [PF8.from_nchw(p) for p in prevs]
    many = [PF8.from_nchw(p) for p in prevs]
    modes = [2, 1, 2, 1]
    for s in range(4):
        tops.fuse_sum_bwd(dyp, y, one[s], s, relu=True, mode=modes[s])
    tops.fuse_sum_bwd_batch(dyp, y, many, [0, 1, 2, 3], modes, relu=True)
    torch.cuda.synchronize()
    for a, b in zip(one, many):
        assert torch.equal(a.buf, b.buf) and b.padding_is_zero()


@pytest.mark.parametrize("align", [True, False])
@pytest.mark.parametrize("sh,sw,dh,dw", [(8, 6, 32, 24), (4, 4, 32, 32), (16, 16, 32, 32), (2, 2, 16, 16)])
def test_bilinear_backward_matches_autograd(align, sh, sw, dh, dw):
    from hrnet_b200 import tops
    from hrnet_b200.ops import PF8
    N, C = 2, 16
    src = _rand(N, C, sh, sw, seed=50).requires_grad_(True)
    F.interpolate(src, size=(dh, dw), mode="bilinear", align_corners=align).backward(_bf16(_rand(N, C, dh, dw, seed=51)))
    dd = PF8.from_nchw(_bf16(_rand(N, C, dh, dw, seed=51)))
    ds = PF8(N, C, sh, sw)
    tops.bilinear_up_bwd(dd, ds, align, mode=1)
    assert _relerr(ds.to_nchw(), src.grad) < 1e-2 and ds.padding_is_zero()


def test_fused_adam_matches_torch():
    from hrnet_b200 import _lib
    from hrnet_b200.flat import FlatParams
    torch.manual_seed(0)
    shapes = [(64, 32, 3, 3), (64,), (64,), (21, 480, 1, 1), (21,), (), (64, 3, 3, 3)]
    params = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.1) for s in shapes]
    params[2].requires_grad_(False)
    ref = [torch.nn.Parameter(p.detach().clone(), requires_grad=p.requires_grad) for p in params]
    opt = torch.optim.Adam([p for p in ref if p.requires_grad], lr=1e-3, weight_decay=1e-4)
    flat = FlatParams(params, conv_meta={0: (64, 32, 32, 9), 3: (21, 480, 480, 1), 6: (64, 27, 32, 1)}, lr=1e-3, weight_decay=1e-4)
    for step in range(3):
        gs = [torch.randn(s, device="cuda") for s in shapes]
        flat.grads.zero_()
        for i, (p, g) in enumerate(zip(ref, gs)):
            if not p.requires_grad:
                continue
            p.grad = g.clone()
            flat.set_grad_from_natural(i, g)
        opt.step()
        flat.adam_step()
        for p, r in zip(params, ref):
            assert torch.allclose(p, r, rtol=1e-5, atol=1e-7), step
        nat = flat.natural_grads()
        for i, (p, g) in enumerate(zip(ref, gs)):
            if p.requires_grad:
                assert torch.equal(nat[i], g)


def test_batched_pack_equals_single_pack():
    from hrnet_b200.ops import ConvLayer, Repacker
    w = _rand(64, 32, 3, 3, seed=60)
    a = ConvLayer(w)
    rp = Repacker(w.device)
    b = ConvLayer(w, repacker=rp)
    pa, _ = a.pack(32, 4)
    pb, _ = b.pack(32, 4)
    pb2, _ = b.pack(64, 2)
    rp.run()
    torch.cuda.synchronize()
    assert torch.equal(pa, pb)
    assert torch.equal(a.pack(64, 2)[0], pb2)
