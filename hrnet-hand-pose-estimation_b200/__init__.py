"""B200-native HRNet hand-pose hot path (drop-in for lib/models/pose_hrnet*.py, lib/utils/heatmap_decoding.py,
lib/core/inference.py and lib/core/loss.py of ZJULiHongxin/HRNet-Hand-Pose-Estimation)."""
__version__ = "0.1.0"
