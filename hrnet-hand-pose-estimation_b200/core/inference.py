"""Drop-in for lib/core/inference.py: numpy in, numpy out, computed on the GPU.

get_max_preds(batch_heatmaps)                      reference :18-46
get_final_preds(config, batch_heatmaps, center, scale)   reference :49-85 (+ utils/transforms.py:50-96)

`*_cuda` variants take/return CUDA tensors and avoid the host round trip.
"""
import numpy as np
import torch

from .. import _lib


def get_max_preds_cuda(hm):
    """hm: CUDA float32 [B,J,h,w] -> (preds [B,J,2], maxvals [B,J,1]) CUDA tensors."""
    B, J, h, w = hm.shape
    hm = hm.contiguous().float()
    preds = torch.empty((B, J, 2), dtype=torch.float32, device=hm.device)
    maxvals = torch.empty((B, J, 1), dtype=torch.float32, device=hm.device)
    with torch.cuda.device(hm.device):
        _lib.check(_lib.lib().hrnb_decode_argmax(hm.data_ptr(), B * J, h, w, 0, 1, preds.data_ptr(),
                                                 maxvals.data_ptr(), None, _lib.stream_ptr()))
    return preds, maxvals


def get_final_preds_cuda(hm, center, scale, post_process=True):
    B, J, h, w = hm.shape
    hm = hm.contiguous().float()
    center = torch.as_tensor(center, dtype=torch.float32, device=hm.device).reshape(B, 2).contiguous()
    scale = torch.as_tensor(scale, dtype=torch.float32, device=hm.device).reshape(B, 2).contiguous()
    preds = torch.empty((B, J, 2), dtype=torch.float32, device=hm.device)
    maxvals = torch.empty((B, J, 1), dtype=torch.float32, device=hm.device)
    with torch.cuda.device(hm.device):
        _lib.check(_lib.lib().hrnb_final_preds(hm.data_ptr(), B, J, h, w, center.data_ptr(), scale.data_ptr(),
                                               int(bool(post_process)), preds.data_ptr(), maxvals.data_ptr(),
                                               _lib.stream_ptr()))
    return preds, maxvals


def get_max_preds(batch_heatmaps):
    assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    hm = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, dtype=np.float32)).cuda(non_blocking=True)
    preds, maxvals = get_max_preds_cuda(hm)
    return preds.cpu().numpy(), maxvals.cpu().numpy()


def get_final_preds(config, batch_heatmaps, center, scale):
    assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    hm = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, dtype=np.float32)).cuda(non_blocking=True)
    preds, maxvals = get_final_preds_cuda(hm, np.asarray(center, np.float32), np.asarray(scale, np.float32),
                                          config.TEST.POST_PROCESS)
    return preds.cpu().numpy(), maxvals.cpu().numpy()
