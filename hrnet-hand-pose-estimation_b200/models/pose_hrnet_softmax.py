"""Drop-in for lib/models/pose_hrnet_softmax.py: `get_pose_net(cfg, is_train, **kwargs)` (reference :563-569)."""
from ._hrnet import PoseHighResolutionNet, build  # noqa: F401


def get_pose_net(cfg, is_train, **kwargs):
    return build(cfg, is_train, "softmax", **kwargs)
