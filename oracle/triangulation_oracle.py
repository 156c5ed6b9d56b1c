"""CPU restatement of the reference's algebraic (DLT) triangulation for SURVEY §8 row (f) / BASELINE configs[3].
TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Follows  lib/utils/misc.py:64-97  (DLT_sii_pytorch: linear system, A^T A + 1e-3 I, shifted inverse iteration from a random unit
vector, two iterations, point = -b_k, homogeneous -> euclidean lib/utils/misc.py:28-35) and its call site
lib/models/triangulation.py:258-261 (one call PER JOINT, each drawing a fresh torch.rand(B, 4, 1) start vector on the CPU).
The reference calls the removed torch.solve(b, A); oracle/make_golden.py runs the unmodified function with
torch.solve := lambda b, A: (torch.linalg.solve(A, b), None) and pins this restatement against it
(tests/golden/triangulation.npz).  fp32 throughout, like the reference (`.float()` at misc.py:82).
"""
import numpy as np
import torch


def start_vectors(B, J, seed):
    """the J start vectors the reference's per-joint loop draws after torch.manual_seed(seed): J calls of
    torch.rand(B, 4, 1) on the CPU generator (NOT one call of J*B*4 values: torch's vectorised path orders large draws
    differently), each normalised to unit length (misc.py:86-88)  ->  float32 [J, B, 4]"""
    torch.manual_seed(seed)
    out = []
    for _ in range(J):
        bk = torch.rand(B, 4, 1).float()
        bk = bk / torch.sqrt(bk.permute(0, 2, 1).matmul(bk))
        out.append(bk.squeeze(-1))
    return torch.stack(out).numpy()


def normal_matrix(points, proj):
    """points [B, V, 2], proj [B, V, 3, 4] -> A^T A [B, 4, 4] (float32), rows u*P[2] - P[0], v*P[2] - P[1] (misc.py:78-82)"""
    points = np.asarray(points, np.float32)
    proj = np.asarray(proj, np.float32)
    A = proj[:, :, 2:3, :] * points[:, :, :, None] - proj[:, :, :2, :]          # [B, V, 2, 4]
    A = A.reshape(A.shape[0], -1, 4)
    return np.einsum("bri,brj->bij", A, A).astype(np.float32)


def dlt_sii(points, proj, bk0, iterations=2):
    """one joint: points [B, V, 2], proj [B, V, 3, 4], bk0 [B, 4] unit start vectors -> euclidean points [B, 3]"""
    Bm = normal_matrix(points, proj) + np.float32(0.001) * np.eye(4, dtype=np.float32)[None]
    bk = np.asarray(bk0, np.float32)[:, :, None]
    for _ in range(iterations):
        bk = np.linalg.solve(Bm, bk).astype(np.float32)
        bk = bk / np.sqrt((bk * bk).sum(axis=1, keepdims=True))
    h = -bk[:, :, 0]
    return h[:, :3] / h[:, 3:4]


def dlt_sii_backward(points, proj, bk0, d_out, iterations=2):
    """adjoint of dlt_sii w.r.t. the points (float64 numpy): the chain the CUDA kernel triangulate_dlt_bwd_kernel follows.
    The reference needs no such function - DLT_sii_pytorch is a torch autograd graph (misc.py:64-97); make_golden.py
    stores the gradients that graph produces and tests/test_oracle_golden.py pins this restatement to them.
    points [B, V, 2], proj [B, V, 3, 4], bk0 [B, 4], d_out [B, 3] -> d_points [B, V, 2]"""
    points, proj = np.asarray(points, np.float64), np.asarray(proj, np.float64)
    A = proj[:, :, 2:3, :] * points[:, :, :, None] - proj[:, :, :2, :]            # [B, V, 2, 4]
    Af = A.reshape(A.shape[0], -1, 4)
    Bm = np.einsum("bri,brj->bij", Af, Af) + 0.001 * np.eye(4)[None]
    bk = np.asarray(bk0, np.float64)
    ys, bs, invn = [], [], []
    for _ in range(iterations):
        y = np.linalg.solve(Bm, bk[:, :, None])[:, :, 0]
        n = 1.0 / np.sqrt((y * y).sum(1, keepdims=True))
        bk = y * n
        ys.append(y); bs.append(bk); invn.append(n)
    g = np.asarray(d_out, np.float64)
    w = bk[:, 3:4]
    db = np.concatenate([g / w, -(g * bk[:, :3]).sum(1, keepdims=True) / (w * w)], axis=1)
    dB = np.zeros_like(Bm)
    for it in reversed(range(iterations)):
        dy = (db - bs[it] * (bs[it] * db).sum(1, keepdims=True)) * invn[it]
        lam = np.linalg.solve(Bm, dy[:, :, None])[:, :, 0]          # B symmetric: B^-T = B^-1
        dB -= lam[:, :, None] * ys[it][:, None, :]
        db = lam
    S = dB + dB.transpose(0, 2, 1)
    dA = np.einsum("bij,bvkj->bvki", S, A)                          # d a = (dB + dB^T) a per row
    return np.einsum("bvki,bvi->bvk", dA, proj[:, :, 2, :]).astype(np.float32)


def triangulate_joints(points, proj, bk0, iterations=2):
    """the per-joint loop of AlgebraicTriangulationNet.forward (triangulation.py:258-261):
    points [B, V, J, 2], proj [B, V, 3, 4], bk0 [J, B, 4] -> [B, J, 3]"""
    points = np.asarray(points, np.float32)
    return np.stack([dlt_sii(points[:, :, k], proj, bk0[k], iterations) for k in range(points.shape[2])], axis=1)


def svd_triangulation(points, proj):
    """lib/utils/misc.py:99-121 (exact smallest singular vector) - the value the two inverse iterations converge to"""
    points = np.asarray(points, np.float64)
    proj = np.asarray(proj, np.float64)
    A = proj[:, :, 2:3, :] * points[:, :, :, None] - proj[:, :, :2, :]
    A = A.reshape(A.shape[0], -1, 4)
    _, _, vh = np.linalg.svd(A)
    h = -vh[:, 3, :]
    return (h[:, :3] / h[:, 3:4]).astype(np.float32)
