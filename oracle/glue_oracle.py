"""CPU restatements of the loop glue around the network.  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Parity status: PINNED - oracle/make_golden.py (glue_fixture) runs the UNMODIFIED reference classes / functions
(HeatmapGenerator, flip_back, the ToTensor + Normalize transforms, GlobalAveragePoolingHead, AlgebraicTriangulationNet's
post-processing) and asserts these restatements reproduce them; tests/golden/glue.npz holds the vectors.

Reference lines followed:
  heatmap_generator      lib/dataset/target_generators/target_generators.py:15-53
  flip_back / flip merge lib/utils/transforms.py:16-30, lib/core/function.py:681-701
  normalize_u8           lib/dataset/transforms/transforms.py:38-51 (torchvision to_tensor + normalize), build.py:82-85
  gap_head               lib/models/pose_hrnet_volumetric.py:22-56
"""
import numpy as np
import torch
import torch.nn.functional as F


def heatmap_generator(joints, res_hw, sigma):
    """joints [J, 3] (u, v, visible) -> float32 [J, h, w]"""
    h, w = res_hw
    size = 6 * sigma + 3
    x = np.arange(0, size, 1, float)
    y = x[:, np.newaxis]
    x0 = y0 = 3 * sigma + 1
    g = np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))
    hms = np.zeros((len(joints), h, w), dtype=np.float32)
    for idx, pt in enumerate(joints):
        if pt[2] > 0:
            px, py = int(pt[0]), int(pt[1])
            if px < 0 or py < 0 or px >= w or py >= h:
                continue
            ul = int(np.round(px - 3 * sigma - 1)), int(np.round(py - 3 * sigma - 1))
            br = int(np.round(px + 3 * sigma + 2)), int(np.round(py + 3 * sigma + 2))
            c, d = max(0, -ul[0]), min(br[0], w) - ul[0]
            a, b = max(0, -ul[1]), min(br[1], h) - ul[1]
            cc, dd = max(0, ul[0]), min(br[0], w)
            aa, bb = max(0, ul[1]), min(br[1], h)
            hms[idx, aa:bb, cc:dd] = np.maximum(hms[idx, aa:bb, cc:dd], g[a:b, c:d])
    return hms


def flip_back(output_flipped, matched_parts):
    out = output_flipped[:, :, :, ::-1].copy()
    for a, b in matched_parts:
        tmp = out[:, a].copy()
        out[:, a] = out[:, b]
        out[:, b] = tmp
    return out


def flip_test_merge(hm, hm_flipped, matched_parts, shift):
    fb = flip_back(hm_flipped, matched_parts)
    if shift:
        fb[:, :, :, 1:] = fb.copy()[:, :, :, 0:-1]
    return (hm + fb) * 0.5


def normalize_u8(img_hwc_u8, mean, std):
    """uint8 [H, W, 3] -> float32 [3, H, W]: to_tensor (/255) then (x - mean) / std, in fp32 like torchvision"""
    x = torch.from_numpy(np.ascontiguousarray(img_hwc_u8)).permute(2, 0, 1).float().div(255)
    m = torch.tensor(mean, dtype=torch.float32).view(3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(3, 1, 1)
    return ((x - m) / s).numpy()


def gap_head(sd, prefix, x):
    """GlobalAveragePoolingHead in eval mode: sd = state_dict, x [B, C, h, w] fp32 -> [B, n_classes]"""
    p = prefix + "."
    for ci, bi in ((0, 1), (4, 5)):
        x = F.conv2d(x, sd[p + "features.%d.weight" % ci], sd[p + "features.%d.bias" % ci], padding=1)
        x = F.batch_norm(x, sd[p + "features.%d.running_mean" % bi], sd[p + "features.%d.running_var" % bi],
                         sd[p + "features.%d.weight" % bi], sd[p + "features.%d.bias" % bi], False, 0.1, 1e-5)
        x = F.relu(F.max_pool2d(x, 2))
    x = x.reshape(x.shape[0], x.shape[1], -1).mean(-1)
    x = F.relu(F.linear(x, sd[p + "head.0.weight"], sd[p + "head.0.bias"]))
    x = F.relu(F.linear(x, sd[p + "head.2.weight"], sd[p + "head.2.bias"]))
    return torch.sigmoid(F.linear(x, sd[p + "head.4.weight"], sd[p + "head.4.bias"]))


def aggregation(fc_weights, inputs, weights=(0.4, 0.2, 0.2, 0.2)):
    """lib/models/multiview_pose_hrnet.py:32-71 - fc_weights: the 12 (= V*(V-1)) [size, size] nn.Linear weights in module order,
    inputs: V tensors [N, C, H, W]; -> V fused tensors (target first, then the other views in order, FC index running)"""
    outs, index = [], 0
    V = len(inputs)
    for i in range(V):
        views = [inputs[i]] + [inputs[j] for j in range(V) if j != i]
        warped = [views[0]]
        for j in range(1, V):
            N, C, H, W = views[j].shape
            warped.append(F.linear(views[j].reshape(N * C, H * W), fc_weights[index]).reshape(N, C, H, W))
            index += 1
        t = torch.zeros_like(views[0])
        for v, w in zip(warped, weights):
            t = t + v * w
        outs.append(t)
    return outs
