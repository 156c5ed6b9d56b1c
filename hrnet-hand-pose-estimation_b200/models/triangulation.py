"""Drop-in for AlgebraicTriangulationNet (lib/models/triangulation.py:183-274; SURVEY §8 row f1, BASELINE configs[3]).

forward(images [b, v, 3, H, W], proj_matrices [b, v, 3, 4], orig_img_size=[640, 480]) ->
    (keypoints_3d [b, J, 3], keypoints_2d [b, v, J, 2], heatmaps [b, v, J, h, w], alg_confidences)

Same steps as the reference: the views are folded into the batch, the volumetric backbone runs once on b*v images, the
soft-argmax (or argmax) decode gives heat-map pixel coordinates, they are scaled to the original image size, and the algebraic
triangulation lifts them to 3-D - here ONE hrnb_triangulate_dlt launch for all (sample, joint) pairs instead of the
reference's per-joint Python loop over DLT_sii_pytorch (:258-261), differentiable w.r.t. the 2-D points (hrnb_triangulate_dlt_bwd)
so that a 3-D loss trains stage4 + last_layer as the reference does (:205-215 freeze everything else).

Reference bug kept visible: with MODEL.ALG_CONFIDENCES true the reference unpacks the backbone's THIRD output - the scalar
trainable_temp - as `alg_confidences` (:230 vs pose_hrnet_volumetric.py:634) and fails at the following .view(); that branch
(confidence-weighted SVD triangulation, multiview.triangulate_batch_of_points) cannot run as written and raises here with an
explanation instead.  The unweighted DLT_sii_pytorch branch (:261-264) is the working path and the one mirrored.
"""
import logging

import torch
import torch.nn as nn

from ..utils.heatmap_decoding import get_final_preds
from ..utils.misc import triangulate_joints
from . import pose_hrnet_softmax, pose_hrnet_volumetric

logger = logging.getLogger(__name__)
_BACKBONES = {"pose_hrnet_volumetric": pose_hrnet_volumetric, "pose_hrnet_softmax": pose_hrnet_softmax}


def _get(node, name, default=None):
    try:
        return node[name]
    except (KeyError, AttributeError):
        return default


class AlgebraicTriangulationNet(nn.Module):
    def __init__(self, config, is_train=True):
        super().__init__()
        m = config["MODEL"]
        self.heatmap_softmax = bool(_get(m, "HEATMAP_SOFTMAX", True))
        self.use_alg_confidences = bool(_get(m, "ALG_CONFIDENCES", False))
        name = _get(m, "BACKBONE_NAME", "pose_hrnet_volumetric")
        if name not in _BACKBONES:
            raise ValueError("BACKBONE_NAME %r: the B200 path provides %s" % (name, sorted(_BACKBONES)))
        self.backbone = _BACKBONES[name].get_pose_net(config, is_train=True)
        path = _get(m, "BACKBONE_MODEL_PATH", "")
        if path:
            ckpt = torch.load(path, map_location="cpu")
            sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
            sd = {k.replace("module.", ""): v for k, v in sd.items()}
            self.backbone.load_state_dict(sd, strict=False)
        # freeze lower layers (reference :205-215)
        for p in self.backbone.parameters():
            p.requires_grad = False
        for p in self.backbone.stage4.parameters():
            p.requires_grad = True
        for p in self.backbone.last_layer.parameters():
            p.requires_grad = True

    @classmethod
    def from_widths(cls, width=32, image_size=(256, 256), device="cuda", num_joints=21):
        """benchmark / test helper: the MHP algebraic-triangulation network with random-init weights"""
        from ..config import make_cfg
        cfg = make_cfg(width, num_joints=num_joints, image_size=image_size, softmax=True, trainable_softmax=True)
        cfg.MODEL["BACKBONE_NAME"] = "pose_hrnet_volumetric"
        cfg.MODEL["ALG_CONFIDENCES"] = False
        cfg.MODEL["VOL_CONFIDENCES"] = False
        cfg.MODEL["BACKBONE_MODEL_PATH"] = ""
        torch.manual_seed(0)
        return cls(cfg, is_train=False).to(device)

    def forward(self, images, proj_matrices=None, orig_img_size=(640, 480)):
        if proj_matrices is None:
            raise ValueError("proj_matrices [b, v, 3, 4] are required")
        b, v = images.shape[:2]
        if self.use_alg_confidences:
            raise RuntimeError(
                "MODEL.ALG_CONFIDENCES = true: the reference unpacks the backbone's third output (the scalar trainable_temp, "
                "pose_hrnet_volumetric.py:634) as alg_confidences (triangulation.py:230) and fails at the next .view(); this "
                "branch cannot run as written. Set ALG_CONFIDENCES to false (the DLT_sii_pytorch branch, triangulation.py:261).")
        out = self.backbone(images.reshape(-1, *images.shape[2:]))
        heatmaps = out[0]
        keypoints_2d = get_final_preds(heatmaps, use_softmax=self.heatmap_softmax)           # b*v x J x 2, heat-map pixels
        heatmaps = heatmaps.view(b, v, *heatmaps.shape[1:])
        keypoints_2d = keypoints_2d.view(b, v, *keypoints_2d.shape[1:])
        hs = heatmaps.shape[-1]                                                              # reference: one size for both axes
        scale = torch.tensor([orig_img_size[0] / hs, orig_img_size[1] / hs], dtype=keypoints_2d.dtype, device=keypoints_2d.device)
        keypoints_2d = keypoints_2d * scale
        keypoints_3d = triangulate_joints(keypoints_2d, proj_matrices)
        return keypoints_3d, keypoints_2d, heatmaps, None
