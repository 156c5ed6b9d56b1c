import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_train_network import _setup
from hrnet_b200.train import TrainEngine
B, H, W = 2, 128, 128
mode = sys.argv[1] if len(sys.argv) > 1 else "module"
m, cfg, sd, x, gt, xy, vis = _setup("softmax", True, B, H, W)
xs, gts, xys, viss = x.cuda(), gt.cuda(), xy.cuda(), vis.cuda()
if mode == "module":
    m(xs)
    p = m.train_engine().plans[(B, H, W)]
elif mode == "graph":
    p = TrainEngine(m, use_graph=True).forward(xs, True)
else:
    p = TrainEngine(m, use_graph=False).forward(xs, True)
torch.cuda.synchronize()
m2, *_ = _setup("softmax", True, B, H, W)
e2 = TrainEngine(m2, use_graph=False)
p2 = e2.forward(xs, False)
torch.cuda.synchronize()
print(mode, "logits diff", float((p.out["logits"] - p2.out["logits"]).abs().max()))
n = 0
for k in p.conv_out:
    d = float((p.conv_out[k].buf.float() - p2.conv_out[k].buf.float()).abs().max())
    if d > 0:
        print("first differing conv output:", k, d)
        n += 1
        if n > 3: break
w1 = m.train_engine().flat.data if mode == "module" else None
print("x equal", bool(torch.equal(p.x, p2.x)))
