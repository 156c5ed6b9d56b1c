"""Deterministic synthetic inputs / weight perturbations shared by make_golden.py, tests and bench.  TEST INFRA.

Synthetic data follows SURVEY.md §8(d): images ~ N(0,1) (post-Normalize statistics), GT heat maps are
peak-1 Gaussians sigma=2 at uniform joint coordinates (lib/dataset/target_generators.py:15-53).
"""
import numpy as np
import torch


def perturb_state_dict(sd, seed=123):
    """Make BN statistics / affine parameters non-trivial (default init is mean 0, var 1, gamma 1, beta 0,
    which would leave BN folding untested).  Operates in place, in key order, with its own generator."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd):            # sorted: independent of module registration order
        v = sd[k]
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif ".bn" in k or k.startswith("bn") or _is_bn_affine(k, sd):
            if k.endswith(".weight"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif k.endswith(".bias"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    return sd


def _is_bn_affine(k, sd):
    stem = k.rsplit(".", 1)[0]
    return (stem + ".running_mean") in sd


def contract_state_dict(sd, factor=0.05):
    """Scale the gamma of every residual-closing BatchNorm (BasicBlock bn2, Bottleneck bn3) and of every fuse-layer
    BatchNorm by `factor`: each block becomes x + small * f(x), so the random-init network no longer amplifies a 0.2 %
    bf16 rounding of its activations into O(1) gradient changes (deep BatchNorm networks at initialisation do; see
    DESIGN.md §2).  Gives a WELL-CONDITIONED training case that can be compared with the fp32 reference directly."""
    import re
    for k in sorted(sd):
        if not k.endswith(".weight") or sd[k].dim() != 1:
            continue
        if re.search(r"branches\.\d+\.\d+\.bn2\.weight$", k) or re.search(r"layer1\.\d+\.bn3\.weight$", k) or \
                (".fuse_layers." in k and k.endswith(".1.weight")):
            sd[k].mul_(factor)
    return sd


def sharpen_head(sd, factor=50.0):
    """Second weight set with peaky heat maps (default-init logits are nearly flat)."""
    sd["last_layer.3.weight"].mul_(factor)
    return sd


def images(B, H=256, W=256, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, H, W, generator=g)


def targets(B, J=21, h=64, w=64, seed=2, sigma=2.0):
    """(gt_heatmaps [B,J,h,w], pose2d_gt [B,J,2], visibility [B,J])"""
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


def multiview_joints(P, J=21, seed=4, spread=80.0, noise_px=1.5):
    """ground-truth 3-D joints [B, J, 3] around the origin and their noisy projections [B, V, J, 2] under P [B, V, 3, 4]"""
    g = torch.Generator().manual_seed(seed)
    B = P.shape[0]
    X = torch.randn(B, J, 3, generator=g) * spread
    Xh = torch.cat([X, torch.ones(B, J, 1)], -1)
    uvw = torch.einsum("bvij,bkj->bvki", P, Xh)
    uv = uvw[..., :2] / uvw[..., 2:]
    return X, uv + torch.randn(uv.shape, generator=g) * noise_px


def tensor_checksums(sd, keys):
    return {k: float(sd[k].double().abs().sum()) for k in keys}


def sample(t, limit=4096, stride=37):
    """flattened tensor, strided down when it has more than `limit` elements (keeps golden files small)"""
    f = t.detach().reshape(-1)
    return f if f.numel() <= limit else f[::stride][:limit * 4].contiguous()
