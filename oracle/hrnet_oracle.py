"""CPU restatement (torch fp32 functional ops) of the reference HRNet forward pass.  TEST INFRASTRUCTURE.

Parity status: PINNED - checked against the unmodified reference modules imported from /root/reference
(oracle/make_golden.py writes tests/golden/hrnet_*.npz; tests/test_oracle_golden.py re-checks them).

It consumes a reference-layout state_dict (keys as produced by lib/models/pose_hrnet*.py) and re-derives
every key from the stage tables, so it also checks the product's key mapping.

Reference lines followed (relative to the reference repo):
  stem                      lib/models/pose_hrnet.py:283-291, 512-517
  Bottleneck / BasicBlock   lib/models/pose_hrnet.py:28-98
  transition layers         lib/models/pose_hrnet.py:419-458, 521-546
  HighResolutionModule      lib/models/pose_hrnet.py:187-266
  head (raw logits)         lib/models/pose_hrnet.py:560-568   (bilinear, align_corners=False)
  head (softmax variant)    lib/models/pose_hrnet_softmax.py:499-528 (align_corners=True, softmax*temp)
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


class Arch:
    """Stage tables of MODEL.EXTRA (channels per branch for stages 2..4, modules per stage)."""

    def __init__(self, channels=(32, 64, 128, 256), modules=(1, 4, 3), blocks=4, num_joints=21):
        self.channels = tuple(channels)
        self.modules = tuple(modules)
        self.blocks = blocks
        self.num_joints = num_joints

    @staticmethod
    def from_cfg(cfg):
        ex = cfg["MODEL"]["EXTRA"]
        ch = tuple(ex["STAGE4"]["NUM_CHANNELS"])
        mods = tuple(ex["STAGE%d" % s]["NUM_MODULES"] for s in (2, 3, 4))
        return Arch(ch, mods, ex["STAGE2"]["NUM_BLOCKS"][0], cfg["MODEL"].get("NUM_JOINTS", 21))


W32 = Arch((32, 64, 128, 256))
W48 = Arch((48, 96, 192, 384))


_BN_TRAIN = False     # forward_train() switches the BatchNorms to batch statistics (nn.BatchNorm2d in train mode)


def _bn(sd, key, x):
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"],
                        sd[key + ".bias"], _BN_TRAIN, 0.1, EPS)


_CONV_HOOK = None     # tests only: fn(key, conv_output) -> tensor used instead (see train_oracle.train_step(conv_hook=))


def _conv(sd, key, x, stride=1, pad=0):
    y = F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=pad)
    return _CONV_HOOK(key, y) if _CONV_HOOK is not None else y


def _basic_block(sd, pre, x):
    out = F.relu(_bn(sd, pre + ".bn1", _conv(sd, pre + ".conv1", x, 1, 1)))
    out = _bn(sd, pre + ".bn2", _conv(sd, pre + ".conv2", out, 1, 1))
    return F.relu(out + x)


def _bottleneck(sd, pre, x, has_down):
    out = F.relu(_bn(sd, pre + ".bn1", _conv(sd, pre + ".conv1", x)))
    out = F.relu(_bn(sd, pre + ".bn2", _conv(sd, pre + ".conv2", out, 1, 1)))
    out = _bn(sd, pre + ".bn3", _conv(sd, pre + ".conv3", out))
    res = _bn(sd, pre + ".downsample.1", _conv(sd, pre + ".downsample.0", x)) if has_down else x
    return F.relu(out + res)


def _hr_module(sd, pre, xs, blocks):
    nb = len(xs)
    xs = list(xs)
    for i in range(nb):
        for b in range(blocks):
            xs[i] = _basic_block(sd, "%s.branches.%d.%d" % (pre, i, b), xs[i])
    outs = []
    for i in range(nb):
        y = None
        for j in range(nb):
            if j == i:
                t = xs[j]
            elif j > i:
                fp = "%s.fuse_layers.%d.%d" % (pre, i, j)
                t = _bn(sd, fp + ".1", _conv(sd, fp + ".0", xs[j]))
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
            else:
                t = xs[j]
                for k in range(i - j):
                    fp = "%s.fuse_layers.%d.%d.%d" % (pre, i, j, k)
                    t = _bn(sd, fp + ".1", _conv(sd, fp + ".0", t, 2, 1))
                    if k != i - j - 1:
                        t = F.relu(t)
            y = t if y is None else y + t
        outs.append(F.relu(y))
    return outs


def backbone(sd, x, arch):
    """Returns (stage4 outputs list, stage3 branch-0 output)."""
    x = F.relu(_bn(sd, "bn1", _conv(sd, "conv1", x, 2, 1)))
    x = F.relu(_bn(sd, "bn2", _conv(sd, "conv2", x, 2, 1)))
    for b in range(4):
        x = _bottleneck(sd, "layer1.%d" % b, x, b == 0)
    ch = arch.channels
    # transition1: branch 0 3x3 s1 (256 -> C0), branch 1 3x3 s2 (256 -> C1)
    xs = [F.relu(_bn(sd, "transition1.0.1", _conv(sd, "transition1.0.0", x, 1, 1))),
          F.relu(_bn(sd, "transition1.1.0.1", _conv(sd, "transition1.1.0.0", x, 2, 1)))]
    stage3_b0 = None
    for s, nmod in zip((2, 3, 4), arch.modules):
        nb = s
        if s > 2:
            key = "transition%d.%d.0" % (s - 1, nb - 1)
            xs = list(xs) + [F.relu(_bn(sd, key + ".1", _conv(sd, key + ".0", xs[-1], 2, 1)))]
        for m in range(nmod):
            xs = _hr_module(sd, "stage%d.%d" % (s, m), xs, arch.blocks)
        if s == 3:
            stage3_b0 = xs[0]
    return xs, stage3_b0


def forward(sd, x, arch=W32, variant="softmax"):
    """variant 'raw'     -> (logits, stage3_branch0)                      [pose_hrnet.py:568]
       variant 'softmax' -> (heatmap, concat_feat, temp, logits)          [pose_hrnet_softmax.py:528]
       (logits are returned additionally for testing)."""
    sd = {k: v.detach().float() for k, v in sd.items() if torch.is_tensor(v)}
    with torch.no_grad():
        return _forward(sd, x, arch, variant)


def forward_train(sd, x, arch=W32, variant="softmax"):
    """model.train() forward with autograd enabled: `sd` maps state_dict keys to tensors; parameters that should
    receive gradients are leaf tensors with requires_grad=True, running statistics are updated IN PLACE
    (momentum 0.1, unbiased variance), exactly as nn.BatchNorm2d does in train mode.  Same return as forward()."""
    global _BN_TRAIN
    _BN_TRAIN = True
    try:
        return _forward(sd, x, arch, variant)
    finally:
        _BN_TRAIN = False


def _forward(sd, x, arch, variant):
    xs, s3b0 = backbone(sd, x.float(), arch)
    h, w = xs[0].shape[2:]
    align = variant == "softmax"
    ups = [xs[0]] + [F.interpolate(t, size=(h, w), mode="bilinear", align_corners=align) for t in xs[1:]]
    cat = torch.cat(ups, 1)
    y = F.relu(_bn(sd, "last_layer.1", _conv(sd, "last_layer.0", cat)))
    kf = sd["last_layer.3.weight"].shape[-1]
    logits = _conv(sd, "last_layer.3", y, 1, 1 if kf == 3 else 0)
    if variant == "raw":
        return logits, s3b0
    temp = sd["trainable_temp"]
    B, J = logits.shape[:2]
    heat = F.softmax(logits.reshape(B, J, -1) * temp, dim=2).reshape(logits.shape)
    return heat, cat, temp, logits
