"""Topology of the HRNet pose network as data.

`layer_specs(arch)` lists every parameterised layer in the order the reference constructs them
(lib/models/pose_hrnet.py:279-350: stem, layer1 [downsample first, :398-415], transition1, stage2,
transition2, stage3, transition3, stage4, last_layer), each with the state_dict prefix the reference's
module tree gives it.  Both the nn.Module shell (models/_hrnet.py) and the CUDA engine (engine.py) are
generated from this table, so state_dict keys cannot drift apart.
"""
from collections import namedtuple

Conv = namedtuple("Conv", "key cin cout k stride bias")
BN = namedtuple("BN", "key ch")


class Arch:
    def __init__(self, channels, modules, blocks, num_joints, final_kernel):
        self.channels = tuple(int(c) for c in channels)      # per-branch channels of stage 4
        self.modules = tuple(int(m) for m in modules)        # NUM_MODULES of stages 2, 3, 4
        self.blocks = int(blocks)                            # BasicBlocks per branch per module
        self.num_joints = int(num_joints)
        self.final_kernel = int(final_kernel)
        if len(self.channels) != 4 or len(self.modules) != 3:
            raise ValueError("HRNet pose nets have 4 branches / 3 multi-resolution stages")

    @property
    def head_channels(self):
        return sum(self.channels)


def _get(node, name):
    return node[name] if isinstance(node, dict) else getattr(node, name)


def arch_from_cfg(cfg):
    """Reads the same cfg keys as the reference (MODEL.EXTRA.STAGE{2,3,4}.*, FINAL_CONV_KERNEL, NUM_JOINTS);
    cfg may be a yacs CfgNode, an attribute dict or a plain nested dict."""
    model = _get(cfg, "MODEL")
    extra = _get(model, "EXTRA")
    stages = [_get(extra, "STAGE%d" % s) for s in (2, 3, 4)]
    for s, st in zip((2, 3, 4), stages):
        nb, nblk, nch = _get(st, "NUM_BRANCHES"), _get(st, "NUM_BLOCKS"), _get(st, "NUM_CHANNELS")
        if nb != len(nblk):
            raise ValueError("NUM_BRANCHES({}) <> NUM_BLOCKS({})".format(nb, len(nblk)))
        if nb != len(nch):
            raise ValueError("NUM_BRANCHES({}) <> NUM_CHANNELS({})".format(nb, len(nch)))
        if nb != s:
            raise ValueError("stage %d must have %d branches" % (s, s))
        if _get(st, "BLOCK") != "BASIC":
            raise ValueError("only BASIC blocks are supported in stages 2-4")
        if _get(st, "FUSE_METHOD") != "SUM":
            raise ValueError("only FUSE_METHOD SUM is supported")
        if len(set(nblk)) != 1:
            raise ValueError("all branches must have the same NUM_BLOCKS")
    ch4 = list(_get(stages[2], "NUM_CHANNELS"))
    for st in stages[:2]:
        c = list(_get(st, "NUM_CHANNELS"))
        if c != ch4[:len(c)]:
            raise ValueError("stage channel tables must be prefixes of STAGE4.NUM_CHANNELS")
    if any(int(c) % 16 for c in ch4):
        raise ValueError("NUM_CHANNELS %s: the sm_100a kernels need every branch width to be a multiple of 16 (K step of the "
                         "bf16 tensor-core instruction); HRNet-W32 / W48 / W64 qualify, W18 (18/36/72/144) does not" % (ch4,))
    try:
        nj = _get(model, "NUM_JOINTS")
    except (KeyError, AttributeError):
        nj = 21                                              # lib/config/default.py:52
    return Arch(ch4, [_get(st, "NUM_MODULES") for st in stages], _get(stages[0], "NUM_BLOCKS")[0], nj,
                _get(extra, "FINAL_CONV_KERNEL"))


def bottleneck_specs(prefix, cin, planes, downsample):
    out = []
    if downsample:   # constructed before the block itself (pose_hrnet.py:399-410)
        out += [Conv(prefix + ".downsample.0", cin, planes * 4, 1, 1, False), BN(prefix + ".downsample.1", planes * 4)]
    out += [Conv(prefix + ".conv1", cin, planes, 1, 1, False), BN(prefix + ".bn1", planes),
            Conv(prefix + ".conv2", planes, planes, 3, 1, False), BN(prefix + ".bn2", planes),
            Conv(prefix + ".conv3", planes, planes * 4, 1, 1, False), BN(prefix + ".bn3", planes * 4)]
    return out


def basic_specs(prefix, ch):
    return [Conv(prefix + ".conv1", ch, ch, 3, 1, False), BN(prefix + ".bn1", ch),
            Conv(prefix + ".conv2", ch, ch, 3, 1, False), BN(prefix + ".bn2", ch)]


def fuse_specs(prefix, ch):
    """fuse_layers.{i}.{j}: j>i 1x1 conv+BN (then nearest up); j<i chain of (i-j) 3x3 s2 conv+BN."""
    out = []
    nb = len(ch)
    for i in range(nb):
        for j in range(nb):
            if j > i:
                out += [Conv("%s.%d.%d.0" % (prefix, i, j), ch[j], ch[i], 1, 1, False), BN("%s.%d.%d.1" % (prefix, i, j), ch[i])]
            elif j < i:
                for k in range(i - j):
                    co = ch[i] if k == i - j - 1 else ch[j]
                    out += [Conv("%s.%d.%d.%d.0" % (prefix, i, j, k), ch[j], co, 3, 2, False),
                            BN("%s.%d.%d.%d.1" % (prefix, i, j, k), co)]
    return out


def layer_specs(arch):
    ch = arch.channels
    specs = [Conv("conv1", 3, 64, 3, 2, False), BN("bn1", 64), Conv("conv2", 64, 64, 3, 2, False), BN("bn2", 64)]
    for b in range(4):
        specs += bottleneck_specs("layer1.%d" % b, 64 if b == 0 else 256, 64, b == 0)
    specs += [Conv("transition1.0.0", 256, ch[0], 3, 1, False), BN("transition1.0.1", ch[0]),
              Conv("transition1.1.0.0", 256, ch[1], 3, 2, False), BN("transition1.1.0.1", ch[1])]
    for s, nmod in zip((2, 3, 4), arch.modules):
        nb = s
        if s > 2:
            t = "transition%d.%d.0" % (s - 1, nb - 1)
            specs += [Conv(t + ".0", ch[nb - 2], ch[nb - 1], 3, 2, False), BN(t + ".1", ch[nb - 1])]
        for m in range(nmod):
            pre = "stage%d.%d" % (s, m)
            for i in range(nb):
                for b in range(arch.blocks):
                    specs += basic_specs("%s.branches.%d.%d" % (pre, i, b), ch[i])
            specs += fuse_specs(pre + ".fuse_layers", ch[:nb])
    hc = arch.head_channels
    specs += [Conv("last_layer.0", hc, hc, 1, 1, True), BN("last_layer.1", hc),
              Conv("last_layer.3", hc, arch.num_joints, arch.final_kernel, 1, True)]
    return specs
