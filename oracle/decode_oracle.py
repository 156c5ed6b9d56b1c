"""numpy restatement of the reference heat-map decoders.  TEST INFRASTRUCTURE.

Parity status: PINNED against the reference functions imported from /root/reference
(oracle/make_golden.py -> tests/golden/decode_*.npz), except spatial_expectation2d, whose kornia version
the reference does not pin ("parity unpinned" at that third-party boundary: restated from kornia's
published dsnt.spatial_expectation2d; anchored on the call site lib/utils/heatmap_decoding.py:100).

  get_max_preds          lib/core/inference.py:18-46
  argmax_hstride         lib/utils/heatmap_decoding.py:102-107   (uses H as the row stride)
  spatial_expectation2d  kornia.geometry.subpix (call site lib/utils/heatmap_decoding.py:100)
  final_preds            lib/core/inference.py:49-85 + lib/utils/transforms.py:50-96
  spatial_softmax        lib/models/pose_hrnet_softmax.py:521-524
"""
import math

import numpy as np


def get_max_preds(hm):
    assert isinstance(hm, np.ndarray) and hm.ndim == 4
    B, J, h, w = hm.shape
    flat = hm.reshape(B, J, -1)
    idx = np.argmax(flat, 2)
    maxvals = np.amax(flat, 2).reshape(B, J, 1)
    preds = np.zeros((B, J, 2), np.float32)
    preds[:, :, 0] = idx % w
    preds[:, :, 1] = np.floor(idx / w)
    preds *= (maxvals > 0.0).astype(np.float32)
    return preds, maxvals


def argmax_hstride(hm):
    B, J, h, w = hm.shape
    idx = np.argmax(hm.reshape(B, J, -1), 2)
    return np.stack((idx % h, idx // h), 2).astype(np.float32)


def spatial_softmax(logits, temp=1.0):
    B, J, h, w = logits.shape
    z = logits.reshape(B, J, -1).astype(np.float32) * np.float32(temp)
    z = z - z.max(2, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(2, keepdims=True)).reshape(B, J, h, w).astype(np.float32)


def spatial_expectation2d(p):
    B, J, h, w = p.shape
    xs = np.arange(w, dtype=np.float32)[None, None, None, :]
    ys = np.arange(h, dtype=np.float32)[None, None, :, None]
    ex = (p * xs).reshape(B, J, -1).sum(-1, dtype=np.float32)
    ey = (p * ys).reshape(B, J, -1).sum(-1, dtype=np.float32)
    return np.stack((ex, ey), -1).astype(np.float32)


def _third_point(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def inverse_affine(center, scale, out_w, out_h):
    """2x3 matrix mapping heat-map coords to image coords (rot = 0), solved from the same three
    point pairs the reference hands to cv2.getAffineTransform (utils/transforms.py:58-93)."""
    scale = np.asarray(scale, np.float32)
    center = np.asarray(center, np.float32)
    src_w = scale[0] * 200.0
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src[0] = center
    src[1] = center + np.array([0, src_w * -0.5], np.float32)
    dst[0] = [out_w * 0.5, out_h * 0.5]
    dst[1] = np.array([out_w * 0.5, out_h * 0.5], np.float32) + np.array([0, out_w * -0.5], np.float32)
    src[2] = _third_point(src[0], src[1])
    dst[2] = _third_point(dst[0], dst[1])
    A = np.concatenate([dst.astype(np.float64), np.ones((3, 1))], 1)   # dst -> src (inv=1)
    return np.linalg.solve(A, src.astype(np.float64)).T                # 2x3


def final_preds(hm, center, scale, post_process=True):
    coords, maxvals = get_max_preds(hm)
    B, J, h, w = hm.shape
    if post_process:
        for n in range(B):
            for p in range(J):
                m = hm[n, p]
                px = int(math.floor(coords[n, p, 0] + 0.5))
                py = int(math.floor(coords[n, p, 1] + 0.5))
                if 1 < px < w - 1 and 1 < py < h - 1:
                    diff = np.array([m[py, px + 1] - m[py, px - 1], m[py + 1, px] - m[py - 1, px]])
                    coords[n, p] += np.sign(diff) * 0.25
    preds = coords.copy()
    for n in range(B):
        t = inverse_affine(center[n], scale[n], w, h)
        for p in range(J):
            preds[n, p] = t @ np.array([coords[n, p, 0], coords[n, p, 1], 1.0])
    return preds, maxvals
