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


def _materialise(root, specs, before=None):
    """before: {spec key: callable} run right before that layer is created (subclasses that register extra modules at the
    point of the construction order - and of the seeded RNG stream - where the reference creates them)"""
    for sp in specs:
        if before and sp.key in before:
            before[sp.key]()
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
    # The reference builds `downsample` BEFORE the Bottleneck (pose_hrnet.py:399-410: same RNG draw order as the spec table)
    # but the block registers it LAST (pose_hrnet.py:66-76): named_parameters() / modules() order - what an index-keyed
    # optimizer.state_dict() and init_weights' walk depend on - must follow the registration order.
    for blk in root._modules["layer1"]._modules.values():
        if "downsample" in blk._modules:
            blk._modules["downsample"] = blk._modules.pop("downsample")      # re-insert = move to the end


class _TrainForward(torch.autograd.Function):
    """model.train() forward/backward through the CUDA training engine (train.TrainEngine): forward runs the
    batch-statistics network, backward the hand-written data-/weight-gradient kernels; parameter gradients come back
    in the parameters' own layout so that any torch optimizer (the reference builds Adam, lib/utils/utils.py:71-92)
    and loss (core/loss.py) can be used unchanged."""

    @staticmethod
    def forward(ctx, model, x, *params):
        eng = model.train_engine()
        want = model.return_features
        p = eng.forward(x, want_features=want, feat_grad=bool(want and model.feat_requires_grad))
        ctx.model, ctx.plan, ctx.generation = model, p, p.generation
        ctx.set_materialize_grads(False)
        out = p.out["heatmap"] if model.variant == "softmax" else p.out["logits"]
        feat = p.feat.clone() if want else x.new_zeros(())
        if not want:
            ctx.mark_non_differentiable(feat)
        return out.clone(), feat

    @staticmethod
    def backward(ctx, d_out, d_feat):
        model, p = ctx.model, ctx.plan
        eng = model.train_engine()
        if p.generation != ctx.generation:
            # one TrainPlan per (batch, H, W) holds the activations of the LAST train-mode forward at that shape; a second
            # forward (siamese / multi-view / consistency losses, or a no_grad train-mode forward in between) overwrote them
            raise RuntimeError(
                "the B200 training engine keeps one set of saved activations per input shape: this backward belongs to "
                "forward #%d of that shape, but forward #%d has run since and overwrote them. Run backward() before the "
                "next train-mode forward of the same shape (or concatenate the inputs into one batch); see INTEGRATION.md."
                % (ctx.generation, p.generation))
        if d_out is None:
            d_out = torch.zeros_like(p.out["logits"])
        with torch.cuda.device(eng.device):
            if model.variant == "softmax":
                p.d_heat.copy_(d_out)
                p.d_coords.zero_()       # the soft-argmax is its own autograd node on the returned heat map
            else:
                p.d_logits.copy_(d_out)
            eng.backward(p)
            grads = eng.flat.natural_grads(fresh=True)
        out = [g if prm.requires_grad else None for g, prm in zip(grads, eng.flat.params)]
        return (None, None) + tuple(out)


class PoseHighResolutionNet(nn.Module):
    """variant 'raw'     : forward -> (logits, stage3_branch0)                  [pose_hrnet.py:568]
       variant 'softmax' : forward -> (heatmap, concat_feat, trainable_temp)    [pose_hrnet_softmax.py:528]"""

    def __init__(self, cfg, variant="raw", before=None, **kwargs):
        super().__init__()
        self.variant = variant
        self.arch = A.arch_from_cfg(cfg)
        self.specs = A.layer_specs(self.arch)
        _materialise(self, self.specs, before)
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
        self._train_engine = None
        self._train_key = None
        # nn.DataParallel (the reference's default wrapper, tools/train.py:250-254) replicates the module on every forward and
        # calls the replicas from one thread per GPU: replicas share this dict by reference (replicate() copies __dict__
        # shallowly), so every device keeps ONE engine across calls, keyed by the ORIGINAL module's parameter versions
        self._shared = {"orig": self, "engines": {}}
        self._train_epoch = 0         # bumped by the training engine on every train-mode forward / step (stale-fold guard)
        self.return_features = True   # set False to skip materialising the NCHW fp32 feature output
        self.feat_requires_grad = False   # True: loss.backward() may flow through the returned feature tensor (train mode)
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
        self._shared["engines"].clear()
        self._engine = None
        self._engine_key = None
        self._tensors = None
        self._train_engine = None
        self._train_key = None

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.invalidate()
        return super().load_state_dict(*a, **k)

    def _param_versions(self):
        if self._tensors is None:
            self._tensors = list(self.parameters()) + list(self.buffers())
        return sum(t._version for t in self._tensors), len(self._tensors), self._train_epoch

    def engine_parameters(self):
        """[(name, parameter)] the CUDA engines own, in named_parameters() order (subclasses with add-on heads exclude theirs)"""
        return list(self.named_parameters())

    def _weights_version(self):
        """version counter of the parameters only (buffers are updated by the training kernels themselves)"""
        return sum(p._version for p in self.parameters())

    def engine(self):
        from ..engine import HRNetEngine
        if getattr(self, "_is_replica", False):
            return self._replica_engine()
        key = self._param_versions()
        if self._engine is None or self._engine_key != key:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("the B200 HRNet runs on CUDA only (no CPU fallback): call .cuda() first")
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            self._engine = HRNetEngine(sd, self.arch, self.variant, dev)
            self._engine_key = key
        return self._engine

    def _replica_engine(self):
        """engine of a DataParallel replica: cached per device on the original module (self._shared), rebuilt when the
        original's parameters changed; built from THIS replica's (broadcast) tensors, on this replica's device"""
        from ..engine import HRNetEngine
        orig = self._shared["orig"]
        key = orig._param_versions() if orig is not None else None
        # replicas hold their (broadcast, non-leaf) parameters in _former_parameters; .parameters() / state_dict() are empty
        sd = {}
        for mod_name, mod in self.named_modules():
            prefix = mod_name + "." if mod_name else ""
            for k, v in getattr(mod, "_former_parameters", {}).items():
                sd[prefix + k] = v.detach()
            for k, v in mod._parameters.items():
                if v is not None:
                    sd[prefix + k] = v.detach()
            for k, v in mod._buffers.items():
                if v is not None:
                    sd[prefix + k] = v
        dev = sd["conv1.weight"].device
        entry = self._shared["engines"].get(dev)
        if entry is None or entry[0] != key or key is None:
            entry = (key, HRNetEngine(sd, self.arch, self.variant, dev))
            self._shared["engines"][dev] = entry
        return entry[1]

    def train_engine(self, **kw):
        """the CUDA training engine of this module (created on first use; parameters become views of its flat fp32
        buffer).  After parameters were changed by a torch optimizer the packed bf16 weights are refreshed."""
        from ..train import TrainEngine
        if self._train_engine is None:
            if next(self.parameters()).device.type != "cuda":
                raise RuntimeError("the B200 HRNet runs on CUDA only (no CPU fallback): call .cuda() first")
            self._train_engine = TrainEngine(self, **kw)
            self._tensors = None
            self._train_key = self._weights_version()
        elif self._train_key != self._weights_version():
            with torch.cuda.device(self._train_engine.device):
                self._train_engine.repack()
            self._train_key = self._weights_version()
        return self._train_engine

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (no CPU fallback)")
        if self.training:
            if getattr(self, "_is_replica", False):
                raise RuntimeError("training under nn.DataParallel is not supported by the B200 engine (it owns flat parameter / "
                                   "optimizer buffers per process): launch one process per GPU (torchrun), as bench.py does; "
                                   "model.eval() forward under nn.DataParallel is supported")
            params = [p for _, p in self.engine_parameters()]
            out, feat = _TrainForward.apply(self, x, *params)
            feat = feat if self.return_features else None
            if self.variant == "softmax":
                return out, feat, self.trainable_temp
            return out, feat
        eng = self.engine()
        out = eng.forward(x, want_features=self.return_features)
        # the engine's outputs are static CUDA-graph buffers; hand out copies unless told otherwise
        get = (lambda k: out[k]) if self.static_outputs else (lambda k: out[k].clone())
        feat = get("features") if self.return_features else None
        if self.variant == "softmax":
            return get("heatmap"), feat, self.trainable_temp
        return get("logits"), feat


    IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)       # lib/dataset/transforms/build.py:85

    def forward_images(self, images_u8, mean=IMAGENET_MEAN, std=IMAGENET_STD):
        """eval-mode forward on raw uint8 NHWC images [B,H,W,3]: ToTensor + Normalize (lib/dataset/transforms/build.py:82-85)
        are applied inside the stem kernel.  Same return values as forward()."""
        if self.training:
            raise RuntimeError("forward_images is an inference entry point (model.eval())")
        if not images_u8.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (no CPU fallback)")
        out = self.engine().forward_u8(images_u8, mean, std, want_features=self.return_features)
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
