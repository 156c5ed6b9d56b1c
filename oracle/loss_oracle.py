"""numpy restatement of the reference losses and their analytic gradients.  TEST INFRASTRUCTURE.

Parity status: PINNED against lib/core/loss.py imported from /root/reference (values and autograd
gradients; oracle/make_golden.py -> tests/golden/loss_*.npz).

  heatmap_loss   lib/core/loss.py:15-28    sum_hw (p-g)^2 (or |p-g|), mean over B*J
  pose2d_loss    lib/core/loss.py:30-50    sum(||p-g||_2 * vis) / max(1, sum vis)   |  sum(||.||)/J
  total          lib/core/function.py:1334-1344  HEATMAP_LOSS_FACTOR*hm + POSE2D_LOSS_FACTOR*pose2d
"""
import numpy as np


def heatmap_loss(pred, gt, mode="l2"):
    assert pred.shape == gt.shape
    d = pred.astype(np.float32) - gt.astype(np.float32)
    per = (d * d if mode == "l2" else np.abs(d)).sum(-1, dtype=np.float32).sum(-1, dtype=np.float32)
    return np.float32(per.mean(dtype=np.float32))


def heatmap_loss_grad(pred, gt, mode="l2"):
    d = pred.astype(np.float32) - gt.astype(np.float32)
    bj = np.float32(np.prod(pred.shape[:-2]))
    return (2.0 * d / bj if mode == "l2" else np.sign(d) / bj).astype(np.float32)


def pose2d_loss(pred, gt, vis=None):
    dist = np.sqrt(((pred.astype(np.float32) - gt.astype(np.float32)) ** 2).sum(-1, dtype=np.float32))
    if vis is not None:
        return np.float32((dist * vis).sum(dtype=np.float32) / max(np.float32(1), vis.sum(dtype=np.float32)))
    return np.float32(dist.sum(dtype=np.float32) / pred.shape[1])


def pose2d_loss_grad(pred, gt, vis=None):
    d = pred.astype(np.float32) - gt.astype(np.float32)
    dist = np.sqrt((d ** 2).sum(-1, keepdims=True))
    denom = max(np.float32(1), vis.sum(dtype=np.float32)) if vis is not None else np.float32(pred.shape[1])
    v = vis[..., None] if vis is not None else np.float32(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.where(dist > 0, d / dist, 0.0) * v / denom
    return g.astype(np.float32)


def total_loss(hm_loss, p2d_loss, hm_factor=1.0, p2d_factor=0.1):
    return np.float32(hm_factor) * hm_loss + np.float32(p2d_factor) * p2d_loss
