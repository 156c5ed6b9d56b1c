from . import pose_hrnet, pose_hrnet_softmax  # noqa: F401
