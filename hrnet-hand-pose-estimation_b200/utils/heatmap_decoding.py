"""Drop-in for lib/utils/heatmap_decoding.py (`get_final_preds(hms, use_softmax=True)`, reference :87-107).

use_softmax=True : integral soft-argmax on pixel grids (kornia spatial_expectation2d(normalized_coordinates=False),
                   reference :99-101); differentiable w.r.t. hms.
use_softmax=False: argmax with u = idx % H, v = idx // H where H = hms.shape[2] (reference :103-107 - the
                   reference really uses the HEIGHT as row stride; reproduced bit-for-bit).
CUDA tensors only - there is no CPU fallback.
"""
import torch

from .. import _lib


class _SoftArgmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hms):
        B, J, h, w = hms.shape
        x = hms.contiguous().float()
        out = torch.empty((B, J, 2), dtype=torch.float32, device=hms.device)
        with torch.cuda.device(hms.device):
            _lib.check(_lib.lib().hrnb_softargmax(x.data_ptr(), B * J, h, w, out.data_ptr(), _lib.stream_ptr()))
        ctx.shape = (h, w)
        ctx.in_dtype = hms.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        # dE[x]/dp_i = x_i, dE[y]/dp_i = y_i : an outer product with the pixel grids (pure broadcast)
        h, w = ctx.shape
        xs = torch.arange(w, dtype=torch.float32, device=g.device).view(1, 1, 1, w)
        ys = torch.arange(h, dtype=torch.float32, device=g.device).view(1, 1, h, 1)
        gi = g[..., 0, None, None] * xs + g[..., 1, None, None] * ys
        return gi.to(ctx.in_dtype)


def get_final_preds(hms, use_softmax=True):
    assert isinstance(hms, torch.Tensor), 'hms should be torch.Tensor'
    assert hms.ndim == 4, 'Heatmap shape should be 4-ndim'
    if not hms.is_cuda:
        raise RuntimeError("get_final_preds runs on CUDA tensors only (no CPU fallback)")
    if use_softmax:
        return _SoftArgmax.apply(hms)
    B, J, h, w = hms.shape
    x = hms.detach().contiguous().float()
    preds = torch.empty((B, J, 2), dtype=torch.float32, device=hms.device)
    with torch.cuda.device(hms.device):
        _lib.check(_lib.lib().hrnb_decode_argmax(x.data_ptr(), B * J, h, w, 1, 0, preds.data_ptr(), None, None,
                                                 _lib.stream_ptr()))
    return preds
