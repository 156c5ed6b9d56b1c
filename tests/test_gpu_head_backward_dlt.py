"""-m gpu: the separable bilinear backward kernel (default since round 2) and the DLT triangulation kernel.  Both were
staged at the end of round 1 behind xfail markers; all 13 cases passed on the driver's B200 (GPUTEST_r01.json), so they are
ordinary tests now and a regression fails the suite."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf16(t):
    return t.to(torch.bfloat16).float()


def _rand(*shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g)


@pytest.mark.parametrize("align", [True, False])
@pytest.mark.parametrize("N,C,sh,sw,dh,dw", [(2, 16, 8, 6, 32, 24), (2, 16, 16, 16, 32, 32), (3, 8, 2, 2, 16, 16),
                                            (8, 256, 8, 8, 64, 64), (4, 64, 32, 32, 64, 64)])
def test_separable_bilinear_backward(align, N, C, sh, sw, dh, dw):
    """hrnb_bilinear_up_bwd (default: one block per image x plane, separable two-pass reduction in shared memory)
    against autograd of F.interpolate and against the gather kernel (hrnb_debug_set(7, 1)), write and accumulate mode"""
    from hrnet_b200 import _lib, tops
    from hrnet_b200.ops import PF8
    src = _rand(N, C, sh, sw, seed=50).requires_grad_(True)
    dy = _bf16(_rand(N, C, dh, dw, seed=51))
    F.interpolate(src, size=(dh, dw), mode="bilinear", align_corners=align).backward(dy)
    dd = PF8.from_nchw(dy)
    ref_kernel = PF8(N, C, sh, sw)
    lib = _lib.lib()
    lib.hrnb_debug_set(7, 1)
    try:
        tops.bilinear_up_bwd(dd, ref_kernel, align, mode=1)
        torch.cuda.synchronize()
    finally:
        lib.hrnb_debug_set(7, 0)
    prev = _bf16(_rand(N, C, sh, sw, seed=52))
    ds = PF8(N, C, sh, sw)
    ds.buf.fill_(3.0)
    ds.buf[:, :ds.lead] = 0
    ds.buf[:, ds.lead + ds.P:] = 0
    tops.bilinear_up_bwd(dd, ds, align, mode=1)
    acc = PF8.from_nchw(prev)
    tops.bilinear_up_bwd(dd, acc, align, mode=2)
    torch.cuda.synchronize()
    scale = max(1e-12, src.grad.abs().max().item())
    assert ds.padding_is_zero() and acc.padding_is_zero()
    assert (ds.to_nchw() - src.grad).abs().max().item() < 1e-2 * scale
    assert (ds.to_nchw() - ref_kernel.to_nchw()).abs().max().item() < 1e-2 * scale
    assert (acc.to_nchw() - (src.grad + prev)).abs().max().item() < 1.5e-2 * max(scale, prev.abs().max().item())


@pytest.mark.parametrize("case", ["mhp4", "two_views", "eight_views_j20"])
def test_dlt_triangulation_matches_reference_golden(case):
    """hrnb_triangulate_dlt (one launch for all joints) against the values of the UNMODIFIED reference DLT_sii_pytorch called
    per joint (tests/golden/triangulation.npz, seeded start vectors) and against the oracle; fp32 tolerance 1e-4 of the
    coordinate scale (the numpy emulation of the kernel's arithmetic sits at 4e-6)."""
    import os
    import numpy as np
    from hrnet_b200.utils import misc
    from oracle import triangulation_oracle as T
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "triangulation.npz"))
    P, uv, bk0, ref = (torch.from_numpy(g[case + "/" + k]).cuda() for k in ("proj", "points", "bk0", "ref"))
    out = misc.triangulate_joints(uv, P, start_vectors=bk0)
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() < 1e-4 * scale
    ours = T.triangulate_joints(uv.cpu().numpy(), P.cpu().numpy(), bk0.cpu().numpy())
    assert np.abs(out.cpu().numpy() - ours).max() < 1e-4 * scale
    # the drop-in entry points draw the reference's random start vectors themselves
    torch.manual_seed(int(g[case + "/seed"]))
    again = misc.triangulate_joints(uv, P)
    assert (again - ref).abs().max().item() < 1e-4 * scale
    torch.manual_seed(int(g[case + "/seed"]))
    one = misc.DLT_sii_pytorch(uv[:, :, 0], P)
    assert (one - ref[:, 0]).abs().max().item() < 1e-4 * scale


@pytest.mark.parametrize("case", ["mhp4", "two_views", "eight_views_j20"])
def test_dlt_triangulation_backward_matches_reference_autograd(case):
    """gradients w.r.t. the 2-D points through hrnb_triangulate_dlt_bwd against the gradients the UNMODIFIED reference's
    autograd produced through DLT_sii_pytorch (tests/golden/triangulation.npz: d_out, d_points)"""
    import os
    import numpy as np
    from hrnet_b200.utils import misc
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "triangulation.npz"))
    P, uv, bk0, d_out, ref = (torch.from_numpy(g[case + "/" + k]).cuda() for k in ("proj", "points", "bk0", "d_out", "d_points"))
    uv = uv.clone().requires_grad_(True)
    out = misc.triangulate_joints(uv, P, start_vectors=bk0)
    (out * d_out).sum().backward()
    scale = ref.abs().max().item()
    assert (uv.grad - ref).abs().max().item() < 2e-3 * scale
    with pytest.raises(NotImplementedError):
        misc.triangulate_joints(uv.detach(), P.clone().requires_grad_(True), start_vectors=bk0)
