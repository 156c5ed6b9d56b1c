"""Import the UNMODIFIED reference modules from /root/reference (authoring container only).  TEST INFRASTRUCTURE.

/root/reference does not exist on the GPU box: nothing run there may import this file.  It is used by
oracle/make_golden.py to generate tests/golden/*.npz and by tests that are skipped when the reference
is absent.

Shims (none of them touches a reference file):
  - numpy.int alias            (lib/models/pose_hrnet.py:331 uses np.int, removed in numpy >= 1.24)
  - stub matplotlib.pyplot     (imported, unused on this path: lib/utils/heatmap_decoding.py:6)
  - kornia.geometry.subpix.spatial_expectation2d restated (kornia is not installed; version unpinned by
    the reference) - pixel-grid expectation without renormalisation
  - yacs-free cfg: PyYAML dict with attribute access
"""
import os
import sys
import types

import numpy as np
import torch
import yaml

VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/build_ref.py (git-ignored copy)


def _resolve_root():
    """HRNB_REFERENCE_ROOT, else /root/reference (authoring container), else the verbatim copy oracle/build_ref.py made
    (the only one present on the GPU box; used by bench.py's reference arm)"""
    env = os.environ.get("HRNB_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/lib/models"):
        return "/root/reference"
    return VENDORED


REF_ROOT = _resolve_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "models"))


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def to_attr(d):
    return AttrDict({k: to_attr(v) for k, v in d.items()}) if isinstance(d, dict) else d


def _spatial_expectation2d(p, normalized_coordinates=True):
    B, C, H, W = p.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=p.dtype, device=p.device),
                            torch.arange(W, dtype=p.dtype, device=p.device), indexing="ij")
    if normalized_coordinates:
        xs, ys = xs / (W - 1) * 2 - 1, ys / (H - 1) * 2 - 1
    f = p.reshape(B, C, -1)
    return torch.cat([(f * xs.reshape(-1)).sum(-1, keepdim=True), (f * ys.reshape(-1)).sum(-1, keepdim=True)], -1)


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    if not hasattr(np, "int"):
        np.int = int
    sys.path.insert(0, os.path.join(REF_ROOT, "lib"))
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot})
    if "kornia" not in sys.modules:
        k, kg, ks = (types.ModuleType(n) for n in ("kornia", "kornia.geometry", "kornia.geometry.subpix"))
        ks.spatial_expectation2d = _spatial_expectation2d
        kg.subpix = ks
        k.geometry = kg
        sys.modules.update({"kornia": k, "kornia.geometry": kg, "kornia.geometry.subpix": ks})
    _installed = True


def load_cfg(rel_yaml="experiments/RHD/RHD_HRNet_w32_softmax_hm-pose2dloss_v1.yaml"):
    with open(os.path.join(REF_ROOT, rel_yaml)) as f:
        cfg = to_attr(yaml.safe_load(f))
    cfg.MODEL.setdefault("NUM_JOINTS", 21)          # default lives in lib/config/default.py:52
    cfg.MODEL.setdefault("TRAINABLE_SOFTMAX", False)
    return cfg


def modules():
    """(pose_hrnet, pose_hrnet_softmax, heatmap_decoding, inference, loss) reference modules."""
    install()
    from models import pose_hrnet, pose_hrnet_softmax
    from utils import heatmap_decoding
    from core import inference, loss
    return pose_hrnet, pose_hrnet_softmax, heatmap_decoding, inference, loss
