import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_train_network import _setup
from hrnet_b200.train import TrainEngine
B, H, W = 2, 128, 128
def fresh():
    m, cfg, sd, x, gt, xy, vis = _setup("softmax", True, B, H, W)
    return m, x.cuda(), gt.cuda(), xy.cuda(), vis.cuda()
m0, xs, gts, xys, viss = fresh()
base = TrainEngine(m0, use_graph=False).train_step(xs, gts, xys, viss, optimizer_step=False).out["logits"].clone()
def rep(tag, lg):
    torch.cuda.synchronize()
    print("%-40s logits diff vs fused-eager %.3e" % (tag, float((lg - base).abs().max())), flush=True)
m1, *_ = fresh(); rep("fused graph train_step", TrainEngine(m1, use_graph=True).train_step(xs, gts, xys, viss, optimizer_step=False).out["logits"])
m2, *_ = fresh(); e2 = TrainEngine(m2, use_graph=False); rep("engine.forward eager feat=0", e2.forward(xs, False).out["logits"])
m3, *_ = fresh(); e3 = TrainEngine(m3, use_graph=False); rep("engine.forward eager feat=1", e3.forward(xs, True).out["logits"])
m4, *_ = fresh(); e4 = TrainEngine(m4, use_graph=True); rep("engine.forward graph feat=1", e4.forward(xs, True).out["logits"])
m5, *_ = fresh(); m5.train_engine(use_graph=False); m5(xs); rep("module eager", m5.train_engine().plans[(B, H, W)].out["logits"])
m6, *_ = fresh(); m6(xs); rep("module graph", m6.train_engine().plans[(B, H, W)].out["logits"])
m7, *_ = fresh()
with torch.no_grad():
    e7 = TrainEngine(m7, use_graph=False)
rep("engine built under no_grad, eager", e7.forward(xs, False).out["logits"])
