"""CPU restatement (torch fp32 autograd) of one reference training step.  TEST INFRASTRUCTURE.

Parity status: PINNED - oracle/make_golden.py runs the UNMODIFIED reference model in train mode with the reference's
own HeatmapLoss / JointsMSELoss / get_final_preds and torch.optim.Adam, asserts this restatement reproduces its
losses, gradients, running statistics and updated parameters, and stores them in tests/golden/train_*.npz.

Reference lines followed (relative to the reference repo):
  forward in train mode          lib/models/pose_hrnet_softmax.py:449-528 / pose_hrnet.py:511-568
  pose2d_pred = get_final_preds  lib/core/function.py:69, lib/utils/heatmap_decoding.py:87-101
  loss weighting                 lib/core/function.py:1334-1344 (HEATMAP_LOSS_FACTOR, POSE2D_LOSS_FACTOR)
  zero_grad / backward / step    lib/core/function.py:101-106
  Adam(lr, weight_decay)         lib/utils/utils.py:71-92
"""
import torch

from . import hrnet_oracle


def _is_param(k, v):
    return v.dtype.is_floating_point and not (k.endswith("running_mean") or k.endswith("running_var"))


def train_step(sd, x, gt_heat, gt_xy, vis, arch, variant, f_hm=1.0, f_p2d=0.1, trainable_temp=False, lr=1e-3,
               weight_decay=1e-4, adam=True, conv_values=None, opt_state=None):
    """sd: reference-layout state_dict (tensors are cloned).  Returns dict(losses=(total, hm, p2d), grads={key: tensor},
    state={key: tensor after the step (parameters after one Adam step, updated running statistics)}, logits=...).

    conv_values ({conv key: tensor}, optional) turns this into the LINEARISED oracle: every conv output is replaced
    in value (not in gradient: straight-through) by the given tensor - the activations a bf16 implementation actually
    produced - so autograd differentiates around THAT forward pass.  The backward pass is linear in the upstream
    gradient once the forward values are fixed, which removes the chaotic forward sensitivity of a deep random-init
    BatchNorm network from a gradient comparison and leaves only the backward arithmetic under test."""
    st = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if _is_param(k, t):
            t = t.float().requires_grad_(k != "trainable_temp" or trainable_temp)
        st[k] = t
    if conv_values is not None:
        hrnet_oracle._CONV_HOOK = lambda key, y: y + (conv_values[key].to(y.dtype) - y).detach()
    try:
        out = hrnet_oracle.forward_train(st, x, arch, variant)
    finally:
        hrnet_oracle._CONV_HOOK = None
    if variant == "softmax":
        heat, _, _, logits = out
        B, J, h, w = heat.shape
        xs = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w)
        ys = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1)
        coords = torch.stack(((heat * xs).sum((2, 3)), (heat * ys).sum((2, 3))), -1)   # kornia spatial_expectation2d
        l_hm = ((heat - gt_heat) ** 2).sum(-1).sum(-1).mean()                          # core/loss.py:15-28
        dist = torch.norm(coords - gt_xy, dim=-1)                                      # core/loss.py:30-50
        l_p2d = (dist * vis).sum() / torch.clamp(vis.sum(), min=1.0)
        total = f_hm * l_hm + f_p2d * l_p2d
    else:
        logits = out[0]
        l_hm = ((logits - gt_heat) ** 2).sum(-1).sum(-1).mean()
        l_p2d = torch.zeros(())
        total = f_hm * l_hm
    leaves = {k: t for k, t in st.items() if torch.is_tensor(t) and t.requires_grad}
    grads = dict(zip(leaves, torch.autograd.grad(total, list(leaves.values()), allow_unused=True)))
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in grads.items()}
    new_opt_state = None
    if adam:
        opt = torch.optim.Adam(list(leaves.values()), lr=lr, weight_decay=weight_decay)
        if opt_state is not None:           # multi-step trajectories: carry the Adam moments / step count over
            opt.load_state_dict(opt_state)
        for k, t in leaves.items():
            t.grad = grads[k].clone()
        opt.step()
        new_opt_state = opt.state_dict()
    state = {k: t.detach() for k, t in st.items()}
    for k in list(state):                   # nn.BatchNorm2d(train) also counts its batches
        if k.endswith("num_batches_tracked"):
            state[k] = state[k] + 1
    return {"losses": (float(total.detach()), float(l_hm.detach()), float(l_p2d.detach())), "grads": grads, "state": state,
            "logits": logits.detach(), "opt_state": new_opt_state}
