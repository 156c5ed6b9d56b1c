"""Device-side drop-ins for the flip-test helpers of lib/utils/transforms.py:16-30 and lib/core/function.py:681-701.

    flip_back(output_flipped [B,J,h,w], matched_parts) -> tensor      (reference: numpy in, numpy out, after a .cpu())
    flip_test_merge(heatmaps, heatmaps_flipped, matched_parts, shift_heatmap) -> 0.5 * (heatmaps + shift(flip_back(flipped)))

The reference moves the flipped prediction to the host, reverses / swaps it in numpy and uploads it again; here it is one kernel
(hrnb_flip_merge) on CUDA tensors.  No CPU fallback."""
import torch

from .. import _lib

_PERMS = {}


def _perm(J, matched_parts, device):
    key = (J, tuple(tuple(p) for p in matched_parts), str(device))
    if key not in _PERMS:
        perm = list(range(J))
        for a, b in matched_parts:          # flip_back swaps channel a and b, pair by pair, in order
            perm[a], perm[b] = perm[b], perm[a]
        _PERMS[key] = torch.tensor(perm, dtype=torch.int32, device=device)
    return _PERMS[key]


def _run(hm, flipped, matched_parts, shift):
    assert flipped.dim() == 4, "output_flipped should be [batch_size, num_joints, height, width]"
    if not flipped.is_cuda:
        raise RuntimeError("the B200 flip helpers take CUDA tensors (no CPU fallback)")
    flipped = flipped.contiguous().float()
    B, J, h, w = flipped.shape
    out = torch.empty_like(flipped)
    perm = _perm(J, matched_parts, flipped.device)
    if hm is not None:
        hm = hm.contiguous().float()
        assert hm.shape == flipped.shape
    with torch.cuda.device(flipped.device):
        _lib.check(_lib.lib().hrnb_flip_merge(hm.data_ptr() if hm is not None else None, flipped.data_ptr(), perm.data_ptr(), B, J,
                                              h, w, int(bool(shift)), out.data_ptr(), _lib.stream_ptr()))
    return out


def flip_back(output_flipped, matched_parts):
    return _run(None, output_flipped, matched_parts, False)


def flip_test_merge(heatmaps, heatmaps_flipped, matched_parts, shift_heatmap=False):
    return _run(heatmaps, heatmaps_flipped, matched_parts, shift_heatmap)
