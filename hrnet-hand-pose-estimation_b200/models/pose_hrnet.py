"""Drop-in for lib/models/pose_hrnet.py: `get_pose_net(cfg, is_train, **kwargs)` (reference :603-609)."""
from ._hrnet import PoseHighResolutionNet, build  # noqa: F401


def get_pose_net(cfg, is_train, **kwargs):
    return build(cfg, is_train, "raw", **kwargs)
