"""Drop-in for lib/models/pose_hrnet_volumetric.py (the backbone AlgebraicTriangulationNet / VolumetricTriangulationNet build,
lib/models/triangulation.py:190,308): the pose_hrnet_softmax network whose forward returns the 4-tuple

    (heatmap [B,J,h,w], inter_feat [B,15*C0,h,w], trainable_temp, vol_confidences [B,32] or None)        reference :634

plus the optional GlobalAveragePoolingHead confidence head on the concat features (reference :22-56, :373-378, :623-625).
State-dict keys, construction (= seeded init) order and cfg keys (MODEL.ALG_CONFIDENCES, MODEL.VOL_CONFIDENCES) follow the
reference.  Reference quirks kept visible instead of silently "fixed":
  * with ALG_CONFIDENCES: true the reference constructor raises NameError (`num_joints` is undefined at :375) - the shipped
    AlgTriangulation_MHP_v1.yaml therefore cannot build its backbone as written; here the head is built with MODEL.NUM_JOINTS
    outputs (state-dict compatible with what the author intended) and, like the reference's forward (:621-622, commented out),
    never evaluated;
  * the confidence head runs in eval mode on the sm_100a kernels (conv + folded BN on the tensor pipe, hrnb_maxpool2_relu,
    hrnb_gap_mlp).  Training THROUGH it belongs to the volumetric pipeline (out of scope) and raises; gradients
    through inter_feat itself are supported (feat_requires_grad).
"""
import torch
import torch.nn as nn

from .. import _lib
from ..ops import PF8, ConvLayer
from ._hrnet import BN_MOMENTUM, PoseHighResolutionNet


class GlobalAveragePoolingHead(nn.Module):
    def __init__(self, in_channels, n_classes):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(in_channels, 512, 3, stride=1, padding=1), nn.BatchNorm2d(512, momentum=BN_MOMENTUM), nn.MaxPool2d(2),
            nn.ReLU(inplace=True),
            nn.Conv2d(512, 256, 3, stride=1, padding=1), nn.BatchNorm2d(256, momentum=BN_MOMENTUM), nn.MaxPool2d(2),
            nn.ReLU(inplace=True))
        self.head = nn.Sequential(nn.Linear(256, 512), nn.ReLU(inplace=True), nn.Linear(512, 256), nn.ReLU(inplace=True),
                                  nn.Linear(256, n_classes), nn.Sigmoid())
        self.n_classes = n_classes
        self._packed = None

    def _layers(self):
        key = sum(t._version for t in list(self.parameters()) + list(self.buffers()))
        if self._packed is None or self._packed[0] != key:
            convs = []
            for ci, bi in ((0, 1), (4, 5)):
                conv, bn = self.features[ci], self.features[bi]
                scale = (bn.weight.detach().float().cpu() / torch.sqrt(bn.running_var.float().cpu() + bn.eps))
                shift = bn.bias.detach().float().cpu() - bn.running_mean.float().cpu() * scale + conv.bias.detach().float().cpu() * scale
                dev = conv.weight.device
                convs.append(ConvLayer(conv.weight.detach().float().contiguous(), scale.to(dev), shift.to(dev), relu=False))
            self._packed = (key, convs)
        return self._packed[1]

    def forward(self, x):
        """x: NCHW fp32 CUDA tensor [B, in_channels, h, w] or a PF8 tensor of the same logical shape -> [B, n_classes]"""
        if self.training:
            raise NotImplementedError("the confidence head is inference-only here (training it is part of the volumetric "
                                      "triangulation pipeline, outside SURVEY §8)")
        if not isinstance(x, PF8):
            if not x.is_cuda:
                raise RuntimeError("GlobalAveragePoolingHead runs on CUDA tensors only (no CPU fallback)")
            x = PF8.from_nchw(x)
        lib = _lib.lib()
        c1, c2 = self._layers()
        dev = x.buf.device
        with torch.cuda.device(dev):
            a = c1(x, PF8(x.N, 512, x.H, x.W, device=dev))
            p1 = PF8(x.N, 512, x.H // 2, x.W // 2, device=dev)
            _lib.check(lib.hrnb_maxpool2_relu(a.ptr, a.ps, a.N, a.C, a.H, a.W, p1.ptr, p1.ps, _lib.stream_ptr()))
            b = c2(p1, PF8(x.N, 256, p1.H, p1.W, device=dev))
            p2 = PF8(x.N, 256, p1.H // 2, p1.W // 2, device=dev)
            _lib.check(lib.hrnb_maxpool2_relu(b.ptr, b.ps, b.N, b.C, b.H, b.W, p2.ptr, p2.ps, _lib.stream_ptr()))
            out = torch.empty((x.N, self.n_classes), dtype=torch.float32, device=dev)
            l1, l2, l3 = self.head[0], self.head[2], self.head[4]
            _lib.check(lib.hrnb_gap_mlp(p2.ptr, p2.ps, p2.N, p2.C, p2.H, p2.W, l1.weight.data_ptr(), l1.bias.data_ptr(), 512,
                                        l2.weight.data_ptr(), l2.bias.data_ptr(), 256, l3.weight.data_ptr(), l3.bias.data_ptr(),
                                        self.n_classes, out.data_ptr(), _lib.stream_ptr()))
        return out


def _flag(model_cfg, name):
    try:
        return bool(model_cfg[name])
    except (KeyError, AttributeError):
        return False


class VolumetricBackbone(PoseHighResolutionNet):
    def __init__(self, cfg, **kwargs):
        m = cfg["MODEL"]

        def add_heads():
            # created where the reference creates them (:373-378): after stage4, before last_layer - same module order and
            # the same position in the seeded RNG stream
            feat = self.arch.head_channels
            if _flag(m, "ALG_CONFIDENCES"):
                self.alg_confidences = GlobalAveragePoolingHead(feat, self.arch.num_joints)
            if _flag(m, "VOL_CONFIDENCES"):
                self.vol_confidences = GlobalAveragePoolingHead(feat, 32)
        super().__init__(cfg, variant="softmax", before={"last_layer.0": add_heads}, **kwargs)
        self.feat_requires_grad = True       # wrappers back-propagate through inter_feat (reference :620-634)

    def engine_parameters(self):
        """the parameters the HRNet engines own (the confidence heads keep theirs)"""
        skip = {id(p) for n in ("alg_confidences", "vol_confidences") if hasattr(self, n) for p in getattr(self, n).parameters()}
        return [(n, p) for n, p in self.named_parameters() if id(p) not in skip]

    def forward(self, x):
        out = super().forward(x)
        heat, feat, temp = out
        vol = None
        if hasattr(self, "vol_confidences"):
            if self.training:
                raise NotImplementedError("training the vol_confidences head is outside SURVEY §8 (volumetric pipeline)")
            vol = self.vol_confidences(feat)
        return heat, feat, temp, vol


def get_pose_net(cfg, is_train, **kwargs):
    """lib/models/pose_hrnet_volumetric.py:669-675"""
    model = VolumetricBackbone(cfg, **kwargs)
    m = cfg["MODEL"]
    init = m["INIT_WEIGHTS"] if "INIT_WEIGHTS" in m else True
    if is_train and init:
        model.init_weights(m["PRETRAINED"] if "PRETRAINED" in m else "")
    return model
