"""Deterministic synthetic workload of the benchmarks: images ~ N(0,1) (post-Normalize statistics,
lib/dataset/transforms/build.py:85), ground-truth heat maps = peak-1 Gaussians (sigma 2) at uniform joint coordinates
(lib/dataset/target_generators.py:15-53), visibility ~ Bernoulli(0.9).  Seeded CPU generators, so every rank / run /
implementation sees the same data (tests/test_host_logic.py checks them against the test fixtures)."""
import torch


def images(B, H=256, W=256, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, H, W, generator=g)


def targets(B, J=21, h=64, w=64, seed=2, sigma=2.0):
    """-> (gt_heatmaps [B,J,h,w], pose2d_gt [B,J,2] as (x, y), visibility [B,J])"""
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(B, J, 2, generator=g) * torch.tensor([w - 1.0, h - 1.0])
    vis = (torch.rand(B, J, generator=g) < 0.9).float()
    ys = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1)
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w)
    mu = xy.round()
    hm = torch.exp(-((xs - mu[..., 0, None, None]) ** 2 + (ys - mu[..., 1, None, None]) ** 2) / (2 * sigma ** 2))
    return hm, xy, vis


def cameras(B, V, seed=3, focal=600.0, center=(320.0, 240.0), distance=600.0):
    """B x V projection matrices K [R | t] of cameras on an arc looking at the origin (MHP-like: 4 views, 640x480)"""
    g = torch.Generator().manual_seed(seed)
    P = torch.zeros(B, V, 3, 4)
    K = torch.tensor([[focal, 0.0, center[0]], [0.0, focal, center[1]], [0.0, 0.0, 1.0]])
    for b in range(B):
        for v in range(V):
            ang = torch.rand(3, generator=g) * 0.6 - 0.3 + torch.tensor([0.0, v * 0.5, 0.0])
            (cx, cy, cz), (sx, sy, sz) = torch.cos(ang), torch.sin(ang)
            Rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
            Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
            Rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
            t = torch.tensor([0.0, 0.0, distance]) + torch.randn(3, generator=g) * 30
            P[b, v] = K @ torch.cat([Rz @ Ry @ Rx, t.view(3, 1)], 1)
    return P
