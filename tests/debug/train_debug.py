"""Debug helper (GPU box): per-parameter gradient parity of one training step against the torch oracle, listed from
the head backwards so the first broken layer of the backward pass is visible.  python tests/debug/train_debug.py [B H W variant]  (test infrastructure: uses the oracle)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fixtures, hrnet_oracle, train_oracle  # noqa: E402
from hrnet_b200.config import make_cfg  # noqa: E402
from hrnet_b200.models import pose_hrnet, pose_hrnet_softmax  # noqa: E402
from hrnet_b200.train import TrainEngine  # noqa: E402

B, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (2, 256, 256)
variant = sys.argv[4] if len(sys.argv) > 4 else "softmax"
cfg = make_cfg(32, softmax=(variant == "softmax"), trainable_softmax=True, image_size=(H, W))
torch.manual_seed(0)
m = (pose_hrnet_softmax if variant == "softmax" else pose_hrnet).get_pose_net(cfg, is_train=False)
sd = m.state_dict()
fixtures.perturb_state_dict(sd)
m.load_state_dict(sd)
sd = {k: v.clone() for k, v in m.state_dict().items()}
x = fixtures.images(B, H, W)
gt, xy, vis = fixtures.targets(B, 21, H // 4, W // 4)
o = train_oracle.train_step(sd, x, gt, xy, vis, hrnet_oracle.Arch.from_cfg(cfg), variant, trainable_temp=True, adam=False)
m = m.cuda().train()
eng = TrainEngine(m, use_graph=False)
p = eng.train_step(x.cuda(), gt.cuda(), xy.cuda(), vis.cuda(), optimizer_step=False)
torch.cuda.synchronize()
print("losses", p.losses.cpu().numpy(), o["losses"])
print("logits rel err", float((p.out["logits"].cpu() - o["logits"]).abs().max() / o["logits"].abs().max()))
names = [n for n, _ in m.named_parameters()]
nat = dict(zip(names, eng.flat.natural_grads()))
gmax = max(float(v.abs().max()) for v in o["grads"].values())
rows = []
for k in names:
    ref = o["grads"].get(k)
    if ref is None:
        continue
    got = nat[k].cpu().double().reshape(-1)
    r = ref.double().reshape(-1)
    l2 = float((got - r).norm() / (r.norm() + 1e-30))
    cos = float((got * r).sum() / (got.norm() * r.norm() + 1e-30))
    rows.append((k, l2, cos, float(r.abs().max()) <= 1e-4 * gmax, float(got.norm()), float(r.norm())))
for k, l2, cos, tiny, gn, rn in reversed(rows):
    flag = "" if (l2 < 0.12 or tiny) else "   <<<<"
    print("%-48s l2 %.4f cos %.4f |got| %.3e |ref| %.3e%s%s" % (k, l2, cos, gn, rn, " (zero-grad)" if tiny else "", flag))
print("---- batch statistics per BN in forward order (rel err of batch mean / var recovered from the running stats)")
bufs = dict(m.named_buffers())
n = 0
for k in sd:
    if not k.endswith("running_mean"):
        continue
    base = k[:-len(".running_mean")]
    for s in ("running_mean", "running_var"):
        old = sd[base + "." + s]
        ref = (o["state"][base + "." + s] - 0.9 * old) / 0.1
        got = (bufs[base + "." + s].cpu() - 0.9 * old) / 0.1
        err = float((got - ref).abs().max() / (ref.abs().max() + 1e-12))
        print("%-44s %-12s rel err %.4f  ref max %.4f%s" % (base, s, err, float(ref.abs().max()), "   <<<<" if err > 0.05 else ""))
    n += 1
    if n >= int(os.environ.get("NBN", "60")):
        break
