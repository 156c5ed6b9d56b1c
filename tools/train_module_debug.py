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
loss = 1.0 * HeatmapLoss()(heat, gts) + 0.1 * JointsMSELoss()(get_final_preds(heat, True), xys, viss)
loss.backward()
eng = m.train_engine()
p = eng.plans[(B, H, W)]
g_auto = {n: q.grad.clone() for n, q in m.named_parameters() if q.grad is not None}
dl_auto = p.d_logits.clone()
m2, *_ = _setup("softmax", True, B, H, W)
eng2 = TrainEngine(m2, use_graph=False)
p2 = eng2.train_step(xs, gts, xys, viss, optimizer_step=False)
torch.cuda.synchronize()
print("loss", float(loss), float(p2.losses[0]))
print("logits diff", float((p.out["logits"] - p2.out["logits"]).abs().max()), "heat diff", float((p.out["heatmap"] - p2.out["heatmap"]).abs().max()))
print("d_logits rel diff", float((dl_auto - p2.d_logits).abs().max() / p2.d_logits.abs().max()))
nat = dict(zip([n for n, _ in m2.named_parameters()], eng2.flat.natural_grads()))
rows = []
for n, g in g_auto.items():
    r = nat[n]
    rows.append((n, float((g - r).norm() / (r.norm() + 1e-30))))
for n, e in list(reversed(rows))[:25] + rows[:8]:
    print("%-50s %.5f" % (n, e))
# second identical fused run on a third model: run-to-run noise of the fused path itself
m3, *_ = _setup("softmax", True, B, H, W)
eng3 = TrainEngine(m3, use_graph=False)
eng3.train_step(xs, gts, xys, viss, optimizer_step=False)
nat3 = dict(zip([n for n, _ in m3.named_parameters()], eng3.flat.natural_grads()))
print("fused vs fused (run-to-run): conv1.weight", float((nat3["conv1.weight"] - nat["conv1.weight"]).norm() / nat["conv1.weight"].norm()),
      "last_layer.3.weight", float((nat3["last_layer.3.weight"] - nat["last_layer.3.weight"]).norm() / nat["last_layer.3.weight"].norm()))
