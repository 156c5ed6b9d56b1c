import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_train_network import _setup
from hrnet_b200.core.loss import HeatmapLoss, JointsMSELoss
from hrnet_b200.utils.heatmap_decoding import get_final_preds
from hrnet_b200.train import TrainEngine
B, H, W = 2, 128, 128
m, cfg, sd, x, gt, xy, vis = _setup("softmax", True, B, H, W)
xs, gts, xys, viss = x.cuda(), gt.cuda(), xy.cuda(), vis.cuda()
heat, feat, temp = m(xs)
p = m.train_engine().plans[(B, H, W)]
torch.cuda.synchronize()
L1 = p.out["logits"].clone(); H1 = heat.clone()
m2, *_ = _setup("softmax", True, B, H, W)
eng2 = TrainEngine(m2, use_graph=False)
p2 = eng2.forward(xs, False)
torch.cuda.synchronize()
first = {k: c.buf.clone() for k, c in p2.conv_out.items()}
stats1 = eng2.stats.clone()
print("m vs m2.first logits", float((L1 - p2.out["logits"]).abs().max()))
n = 0
for k in p.conv_out:
    d = float((p.conv_out[k].buf.float() - first[k].float()).abs().max())
    if d > 0:
        print("  m vs m2.first differ at", k, d, "max", float(first[k].float().abs().max())); n += 1
        if n >= 3: break
p2 = eng2.forward(xs, False)
torch.cuda.synchronize()
print("m2.first vs m2.second logits", float((L1 - p2.out["logits"]).abs().max()), "(vs m)")
n = 0
for k in p2.conv_out:
    d = float((p2.conv_out[k].buf.float() - first[k].float()).abs().max())
    if d > 0:
        print("  m2.first vs m2.second differ at", k, d); n += 1
        if n >= 3: break
print("weights equal m vs m2:", bool(torch.equal(m.train_engine().flat.data, eng2.flat.data)))
print("x equal:", bool(torch.equal(p.x, p2.x)))
