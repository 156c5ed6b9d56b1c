"""Profiling helper: `steps` eager (no CUDA graph, single stream) training steps of HRNet-W32 256x256 at batch B, for
ncu launch lists / full captures:  ncu ... python tools/train_once.py 64 2"""
import os
import sys

os.environ["HRNB_NO_GRAPH"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrnet_b200 import synthetic as fixtures  # noqa: E402
from hrnet_b200.config import make_cfg  # noqa: E402
from hrnet_b200.models import pose_hrnet_softmax  # noqa: E402
from hrnet_b200.train import TrainEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = make_cfg(32)
torch.manual_seed(0)
m = pose_hrnet_softmax.get_pose_net(cfg, is_train=False).cuda().train()
eng = TrainEngine(m, use_graph=False, multi_stream=False)
x = fixtures.images(B, 256, 256).cuda()
gt, xy, vis = (t.cuda() for t in fixtures.targets(B, 21, 64, 64))
for i in range(steps):
    p = eng.train_step(x, gt, xy, vis)
torch.cuda.synchronize()
print("losses", p.losses.tolist(), "launches/step", eng.launches_per_step(p))
