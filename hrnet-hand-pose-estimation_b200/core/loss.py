"""Drop-in for HeatmapLoss and JointsMSELoss of lib/core/loss.py (:15-28, :30-50), forward and backward in
single-pass CUDA kernels.  0-dim tensors with grad, `.cuda()`-able modules, no device->host sync
(the reference's `max(1, tensor)` at :47 forces one; here the clamp happens on the device)."""
import torch
import torch.nn as nn

from .. import _lib


class _HeatmapLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, mode):
        p = pred.contiguous().float()
        g = gt.contiguous().float()
        hw = p.shape[-1] * p.shape[-2]
        bj = p.numel() // hw
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        ws = torch.empty(1024, dtype=torch.float32, device=p.device)
        need = pred.requires_grad
        dpred = torch.empty_like(p) if need else None
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().hrnb_loss_heatmap(p.data_ptr(), g.data_ptr(), bj, hw, mode, loss.data_ptr(),
                                                    dpred.data_ptr() if need else None, None, ws.data_ptr(),
                                                    _lib.stream_ptr()))
        ctx.dpred = dpred
        ctx.in_dtype = pred.dtype
        ctx.in_shape = pred.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.dpred is None:
            return None, None, None
        return (ctx.dpred * g).to(ctx.in_dtype).reshape(ctx.in_shape), None, None


class HeatmapLoss(nn.Module):
    def __init__(self, mode='l2'):
        super().__init__()
        self.mode = mode

    def forward(self, pred, gt):
        assert pred.size() == gt.size(), \
            'Heatmap loss error: prediced heatmaps have size {}, but the groundtruth has {}'.format(pred.shape, gt.shape)
        if self.mode not in ('l2', 'l1'):
            raise ValueError("mode must be 'l2' or 'l1'")
        if not pred.is_cuda:
            raise RuntimeError("HeatmapLoss runs on CUDA tensors only (no CPU fallback)")
        return _HeatmapLossFn.apply(pred, gt, 0 if self.mode == 'l2' else 1)


class _Pose2dLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, vis):
        p = pred.contiguous().float()
        g = gt.contiguous().float()
        v = vis.contiguous().float() if vis is not None else None
        B, J = p.shape[0], p.shape[1]
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        need = pred.requires_grad
        dpred = torch.empty_like(p) if need else None
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().hrnb_loss_pose2d(p.data_ptr(), g.data_ptr(), v.data_ptr() if v is not None else None,
                                                   B, J, loss.data_ptr(), dpred.data_ptr() if need else None,
                                                   _lib.stream_ptr()))
        ctx.dpred = dpred
        ctx.in_dtype = pred.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.dpred is None:
            return None, None, None
        return (ctx.dpred * g).to(ctx.in_dtype), None, None


class JointsMSELoss(nn.Module):
    """Despite the name: visibility-masked mean Euclidean distance of 2-D coordinates (reference :30-50)."""

    def __init__(self):
        super(JointsMSELoss, self).__init__()

    def forward(self, pose2D_pred, pose2D_gt, visibility=None):
        if not pose2D_pred.is_cuda:
            raise RuntimeError("JointsMSELoss runs on CUDA tensors only (no CPU fallback)")
        assert pose2D_pred.shape == pose2D_gt.shape and pose2D_pred.shape[-1] == 2
        return _Pose2dLossFn.apply(pose2D_pred, pose2D_gt, visibility)


def total_loss(heatmap_loss, pose2d_loss, cfg):
    """AverageMeter.computeLosses weighting (lib/core/function.py:1334-1344) without the .item() syncs."""
    return cfg.LOSS.HEATMAP_LOSS_FACTOR * heatmap_loss + cfg.LOSS.POSE2D_LOSS_FACTOR * pose2d_loss
