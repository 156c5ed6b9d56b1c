"""Debug helper (GPU box): poison the caching allocator's free memory with NaNs, then run one forward / step of a fresh
engine: any read of never-written torch.empty memory shows up as NaN."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_train_network import _setup
from hrnet_b200.train import TrainEngine
B, H, W = 2, 128, 128
m, cfg, sd, x, gt, xy, vis = _setup("softmax", True, B, H, W)
xs, gts, xys, viss = x.cuda(), gt.cuda(), xy.cuda(), vis.cuda()
junk = torch.full((1 << 29,), float("nan"), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
del junk
eng = TrainEngine(m, use_graph=False)
p = eng.plan(B, H, W)
p.x.copy_(xs); p.gt_heat.copy_(gts); p.gt_xy.copy_(xys); p.vis.copy_(viss)
for i, fn in enumerate(p.fwd_fns):
    fn(); torch.cuda.synchronize()
    if not bool(torch.isfinite(eng.stats).all()):
        print("first non-finite BN statistics after fwd step", i, p.fwd_names[i])
        break
print("logits finite", bool(torch.isfinite(p.out["logits"]).all()))
p.run_loss(); torch.cuda.synchronize(); print("losses", p.losses.tolist())
for i, fn in enumerate(p.bwd_fns):
    fn(); torch.cuda.synchronize()
    if not bool(torch.isfinite(eng.flat.grads).all()) or not bool(torch.isfinite(eng.stats).all()):
        print("first non-finite gradient after bwd step", i, p.bwd_names[i]); break
print("done")
