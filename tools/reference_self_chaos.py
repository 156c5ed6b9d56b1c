import sys, copy
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from oracle import ref_shim, fixtures
from oracle.make_golden import warm_batch, _ref_losses
pose_hrnet, pose_hrnet_softmax, hd, _, loss = ref_shim.modules()
def run(threads, steps=20):
    torch.set_num_threads(threads)
    cfg = ref_shim.load_cfg("experiments/RHD/RHD_HRNet_w32_softmax_hm-pose2dloss_v1.yaml")
    cfg.MODEL["TRAINABLE_SOFTMAX"] = True
    torch.manual_seed(0)
    ref = pose_hrnet_softmax.get_pose_net(cfg, is_train=False)
    sd0 = copy.deepcopy(ref.state_dict()); fixtures.perturb_state_dict(sd0); ref.load_state_dict(sd0); ref.train()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, ref.parameters()), lr=1e-3, weight_decay=1e-4)
    out=[]
    for s in range(steps):
        x, gt, xy, vis = warm_batch(s, 8, 128, 128)
        t,a,b = _ref_losses(ref, loss, hd, "softmax", x, gt, xy, vis)
        opt.zero_grad(); t.backward(); opt.step()
        out.append([float(t),float(a),float(b)])
    return np.array(out)
a=run(8); b=run(3)
print("rel dev per step (total, hm, p2d):")
print(np.round(np.abs(a-b)/np.abs(a),5))
