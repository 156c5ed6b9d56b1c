"""Debug helper (GPU box): repeat a small training step under different launch modes and report non-finite gradients /
polluted PF8 guards at the end of each step.  python tools/train_race.py B H W"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrnet_b200 import synthetic as fixtures
from hrnet_b200 import _lib
from hrnet_b200.config import make_cfg
from hrnet_b200.models import pose_hrnet_softmax
from hrnet_b200.train import TrainEngine
B, H, W = (int(a) for a in sys.argv[1:4])
cfg = make_cfg(32, softmax=True, trainable_softmax=True, image_size=(H, W))
x = fixtures.images(B, H, W).cuda()
gt, xy, vis = (t.cuda() for t in fixtures.targets(B, 21, H // 4, W // 4))

def trial(tag, no_pdl, sync_each, reps=4):
    _lib.lib().hrnb_debug_set(2, 1 if no_pdl else 0)
    torch.manual_seed(0)
    m = pose_hrnet_softmax.get_pose_net(cfg, is_train=False).cuda().train()
    eng = TrainEngine(m, use_graph=False)
    p = eng.plan(B, H, W)
    names = [n for n, _ in m.named_parameters()]
    for r in range(reps):
        p.x.copy_(x); p.gt_heat.copy_(gt); p.gt_xy.copy_(xy); p.vis.copy_(vis)
        for fn in p.fwd_fns:
            fn()
            if sync_each: torch.cuda.synchronize()
        p.run_loss()
        first = None
        for i, fn in enumerate(p.bwd_fns):
            fn()
            if sync_each: torch.cuda.synchronize()
        torch.cuda.synchronize()
        bad = [n for j, n in enumerate(names) if not bool(torch.isfinite(eng.flat.grad_view(j)).all())]
        pol = [(bi, type(b).__name__, b.C, b.H) for bi, b in enumerate(p.all_bufs) if not b.padding_is_zero()]
        nf = [(bi, type(b).__name__, b.C, b.H) for bi, b in enumerate(p.all_bufs) if not bool(torch.isfinite(b.buf.float()).all())]
        print(tag, "rep", r, "loss", float(p.losses[0]), "nan-grads", len(bad), bad[-1:] , "polluted", len(pol), pol[:3], "nonfinite bufs", len(nf), nf[:3], flush=True)

trial("pdl+async", False, False)
trial("nopdl+async", True, False)
trial("pdl+sync", False, True)
trial("nopdl+sync", True, True)
