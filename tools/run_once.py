"""Run the whole plan eagerly (no CUDA graph, single pass after a warm-up) - the command profiled by ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["HRNB_NO_GRAPH"] = "1"
import torch  # noqa: E402
from hrnet_b200.config import make_cfg  # noqa: E402
from hrnet_b200.models import pose_hrnet_softmax  # noqa: E402
from hrnet_b200 import synthetic as fixtures  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
cfg = make_cfg(32)
torch.manual_seed(0)
model = pose_hrnet_softmax.get_pose_net(cfg, is_train=False).to(dev).eval()
model.return_features = False
model.static_outputs = True
x = fixtures.images(B, 256, 256).to(dev)
for _ in range(passes):
    model(x)
torch.cuda.synchronize()
print("launches per pass", model.engine().plan(B, 256, 256).launches(False))
