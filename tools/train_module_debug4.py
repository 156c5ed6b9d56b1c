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
L2 = p2.out["logits"].clone()
print("1. module fwd vs fused fwd:", float((L1 - L2).abs().max()))
loss = 1.0 * HeatmapLoss()(heat, gts) + 0.1 * JointsMSELoss()(get_final_preds(heat, True), xys, viss)
print("2. after loss: m logits changed by", float((p.out["logits"] - L1).abs().max()), "heat", float((p.out["heatmap"] - H1).abs().max()))
loss.backward()
torch.cuda.synchronize()
print("3. after backward: m logits changed by", float((p.out["logits"] - L1).abs().max()), "m2 logits changed by", float((p2.out["logits"] - L2).abs().max()))
p2 = eng2.train_step(xs, gts, xys, viss, optimizer_step=False)
torch.cuda.synchronize()
print("4. after m2 train_step: m logits changed by", float((p.out["logits"] - L1).abs().max()), "m2 logits vs its fwd-only", float((p2.out["logits"] - L2).abs().max()))
print("losses", float(loss), float(p2.losses[0]), "hm/p2d", p2.losses.tolist())
l_hm = HeatmapLoss()(p2.out["heatmap"], gts); l_p = JointsMSELoss()(get_final_preds(p2.out["heatmap"], True), xys, viss)
print("recomputed via modules on m2 heat:", float(l_hm), float(l_p), "on m heat:", float(HeatmapLoss()(H1, gts)), float(JointsMSELoss()(get_final_preds(H1, True), xys, viss)))
