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
