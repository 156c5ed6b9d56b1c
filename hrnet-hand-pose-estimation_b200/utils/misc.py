"""Drop-in for the algebraic-triangulation functions of lib/utils/misc.py (BASELINE configs[3]).

    DLT_sii_pytorch(points [B,V,2], proj_matricies [B,V,3,4], number_of_iterations=2) -> [B,3]     reference :64-97
    homogeneous_to_euclidean(points [N,M+1]) -> [N,M]                                               reference :28-35
    triangulate_joints(keypoints_2d [B,V,J,2], proj_matrices [B,V,3,4]) -> [B,J,3]
        = the per-joint loop of AlgebraicTriangulationNet.forward, lib/models/triangulation.py:258-261, as ONE launch

Same random behaviour as the reference: every DLT_sii_pytorch call draws its start vector with torch.rand(B, 4, 1) on the
host generator and normalises it (reference :86-88), so a seeded run reproduces the reference's values; triangulate_joints
draws J vectors in joint order, exactly what the reference's loop consumes.  CUDA tensors only - no CPU fallback.
Differentiable like the reference's torch graph (lib/core/function.py train3D back-propagates the 3-D loss through the
triangulation into the backbone): gradients flow to the 2-D points through the hand-written adjoint kernel
(hrnb_triangulate_dlt_bwd); the projection matrices are camera DATA - asking for their gradient raises instead of
silently returning None.  Forward and backward are pinned to the unmodified reference (tests/golden/triangulation.npz).
"""
import torch

from .. import _lib


def homogeneous_to_euclidean(points):
    return (points.transpose(1, 0)[:-1] / points.transpose(1, 0)[-1]).transpose(1, 0)


def _start_vectors(B, J, device):
    out = []
    for _ in range(J):                       # J separate draws, like the reference's per-joint calls
        bk = torch.rand(B, 4, 1).float()
        bk = bk / torch.sqrt(bk.permute(0, 2, 1).matmul(bk))
        out.append(bk.squeeze(-1))
    return torch.stack(out).to(device)


class _DLT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, P, bk0, iterations):
        B, V, J, _ = pts.shape
        out = torch.empty((B, J, 3), dtype=torch.float32, device=pts.device)
        with torch.cuda.device(pts.device):
            _lib.check(_lib.lib().hrnb_triangulate_dlt(pts.data_ptr(), P.data_ptr(), bk0.data_ptr(), B, V, J, int(iterations),
                                                       out.data_ptr(), _lib.stream_ptr()))
        ctx.save_for_backward(pts, P, bk0)
        ctx.iterations = int(iterations)
        return out

    @staticmethod
    def backward(ctx, d_out):
        pts, P, bk0 = ctx.saved_tensors
        B, V, J, _ = pts.shape
        d_pts = torch.empty_like(pts)
        d_out = d_out.contiguous().float()
        with torch.cuda.device(pts.device):
            _lib.check(_lib.lib().hrnb_triangulate_dlt_bwd(pts.data_ptr(), P.data_ptr(), bk0.data_ptr(), d_out.data_ptr(), B, V,
                                                           J, ctx.iterations, d_pts.data_ptr(), _lib.stream_ptr()))
        return d_pts, None, None, None


def _launch(points_bvj2, proj, bk0, iterations):
    if not (points_bvj2.is_cuda and proj.is_cuda):
        raise RuntimeError("the B200 triangulation runs on CUDA tensors only (no CPU fallback)")
    B, V, J, two = points_bvj2.shape
    assert two == 2 and tuple(proj.shape) == (B, V, 3, 4), (tuple(points_bvj2.shape), tuple(proj.shape))
    if proj.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("gradients w.r.t. the projection matrices are not implemented (they are camera data in "
                                  "the reference's pipelines); detach them")
    pts = points_bvj2.contiguous().float()
    P = proj.detach().contiguous().float()
    bk0 = bk0.detach().contiguous().float()
    assert tuple(bk0.shape) == (J, B, 4)
    return _DLT.apply(pts, P, bk0, iterations)


def triangulate_joints(keypoints_2d, proj_matrices, number_of_iterations=2, start_vectors=None):
    """keypoints_2d [B,V,J,2] (image coordinates), proj_matrices [B,V,3,4] -> 3-D joints [B,J,3]"""
    B, V, J, _ = keypoints_2d.shape
    bk0 = start_vectors if start_vectors is not None else _start_vectors(B, J, keypoints_2d.device)
    return _launch(keypoints_2d, proj_matrices, bk0, number_of_iterations)


def DLT_sii_pytorch(points, proj_matricies, number_of_iterations=2):
    """points [B,V,2], proj_matricies [B,V,3,4] -> [B,3] (same name, argument order and spelling as the reference)"""
    B = proj_matricies.shape[0]
    pts = points.reshape(B, -1, 1, 2)
    return _launch(pts, proj_matricies, _start_vectors(B, 1, points.device), number_of_iterations)[:, 0]
