"""nn.Module shell of PoseHighResolutionNet: same constructor contract, attribute names, parameter
initialisation order and state_dict keys as lib/models/pose_hrnet.py / pose_hrnet_softmax.py, but the
module tree is materialised from the layer table in arch.py and `forward` runs the sm_100a engine.

There is no PyTorch/cuDNN forward here: without libhrnb.so (or on a CPU tensor) forward raises.
"""
import logging
import os

import torch
import torch.nn as nn

from .. import arch as A

logger = logging.getLogger(__name__)
BN_MOMENTUM = 0.1


class Node(nn.Module):
    """Generic container: children are registered under the path component the reference uses
    ('0', '1', 'branches', 'conv1', ...); integer indexing mirrors nn.Sequential / nn.ModuleList."""

    def __getitem__(self, idx):
        return self._modules.get(str(idx))

    def __len__(self):
        idx = [int(k) for k in self._modules if k.isdigit()]
        return max(idx) + 1 if idx else 0

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def forward(self, *a, **k):
        raise RuntimeError("sub-modules of the B200 HRNet are parameter holders; call the network's forward")


def _materialise(root, specs):
    for sp in specs:
        parts = sp.key.split(".")
        node = root
        for comp in parts[:-1]:
            nxt = node._modules.get(comp)
            if nxt is None:
                nxt = Node()
                node.add_module(comp, nxt)
            node = nxt
        if isinstance(sp, A.Conv):
            leaf = nn.Conv2d(sp.cin, sp.cout, kernel_size=sp.k, stride=sp.stride, padding=sp.k // 2, bias=sp.bias)
        else:
            leaf = nn.BatchNorm2d(sp.ch, momentum=BN_MOMENTUM)
        node.add_module(parts[-1], leaf)


class PoseHighResolutionNet(nn.Module):
    """variant 'raw'     : forward -> (logits, stage3_branch0)                  [pose_hrnet.py:568]
       variant 'softmax' : forward -> (heatmap, concat_feat, trainable_temp)    [pose_hrnet_softmax.py:528]"""

    def __init__(self, cfg, variant="raw", **kwargs):
        super().__init__()
        self.variant = variant
        self.arch = A.arch_from_cfg(cfg)
        self.specs = A.layer_specs(self.arch)
        _materialise(self, self.specs)
        extra = cfg["MODEL"]["EXTRA"]
        self.pretrained_layers = extra["PRETRAINED_LAYERS"] if "PRETRAINED_LAYERS" in extra else ["*"]
        # keep the reference's index layout: transitionN[i] is None where the branch passes through
        if variant == "softmax":
            model = cfg["MODEL"]
            trainable = bool(model["TRAINABLE_SOFTMAX"]) if "TRAINABLE_SOFTMAX" in model else False
            self.trainable_temp = nn.Parameter(torch.tensor(1.0), requires_grad=trainable)
        self._engine = None
        self._engine_key = None
        self._tensors = None
        self.return_features = True   # set False to skip materialising the NCHW fp32 feature output
        self.static_outputs = False   # True: return the engine's static buffers (overwritten by the next call)

    # ---- reference API -----------------------------------------------------------------------------
    def init_weights(self, pretrained=""):
        """N(0, 0.001) convs, BN gamma=1 beta=0, then partial load (pose_hrnet.py:570-600)."""
        logger.info("=> init weights from normal distribution")
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if pretrained and os.path.isfile(pretrained):
            state = torch.load(pretrained, map_location="cpu")
            logger.info("=> loading pretrained model {}".format(pretrained))
            keep = {k: v for k, v in state.items()
                    if k.split(".")[0] in self.pretrained_layers or self.pretrained_layers[0] == "*"}
            self.load_state_dict(keep, strict=False)
        elif pretrained:
            logger.error("=> please download pre-trained models first!")
            raise ValueError("{} does not exist!".format(pretrained))
        self.invalidate()

    def invalidate(self):
        """Drop the packed weights (call after mutating parameters in place)."""
        self._engine = None
        self._engine_key = None
        self._tensors = None

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.invalidate()
        return super().load_state_dict(*a, **k)

    def _param_versions(self):
        if self._tensors is None:
            self._tensors = list(self.parameters()) + list(self.buffers())
        return sum(t._version for t in self._tensors), len(self._tensors)

    def engine(self):
        from ..engine import HRNetEngine
        key = self._param_versions()
        if self._engine is None or self._engine_key != key:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("the B200 HRNet runs on CUDA only (no CPU fallback): call .cuda() first")
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            self._engine = HRNetEngine(sd, self.arch, self.variant, dev)
            self._engine_key = key
        return self._engine

    def forward(self, x):
        if self.training:
            raise NotImplementedError(
                "training-mode forward/backward (batch-stat BN, dgrad/wgrad kernels) is not built yet; "
                "call model.eval() - there is deliberately no PyTorch/cuDNN fallback")
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (no CPU fallback)")
        eng = self.engine()
        out = eng.forward(x, want_features=self.return_features)
        # the engine's outputs are static CUDA-graph buffers; hand out copies unless told otherwise
        get = (lambda k: out[k]) if self.static_outputs else (lambda k: out[k].clone())
        feat = get("features") if self.return_features else None
        if self.variant == "softmax":
            return get("heatmap"), feat, self.trainable_temp
        return get("logits"), feat


def build(cfg, is_train, variant, **kwargs):
    model = PoseHighResolutionNet(cfg, variant=variant, **kwargs)
    m = cfg["MODEL"]
    init = m["INIT_WEIGHTS"] if "INIT_WEIGHTS" in m else True       # default True: lib/config/default.py:49
    if is_train and init:
        model.init_weights(m["PRETRAINED"] if "PRETRAINED" in m else "")
    return model
