"""yacs-free cfg objects with the keys `get_pose_net` reads (lib/config/default.py + the experiment YAMLs).

`make_cfg(width=32)` reproduces MODEL.* of experiments/RHD/RHD_HRNet_w32_*.yaml (w48: NUM_CHANNELS x1.5);
`load_yaml(path)` loads a reference experiment file.  Objects support attribute AND mapping access, like
the yacs CfgNode the reference passes around (pose_hrnet.py:279 vs :292).
"""


class CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def to_cfg(d):
    return CfgNode({k: to_cfg(v) for k, v in d.items()}) if isinstance(d, dict) else d


def make_cfg(width=32, num_joints=21, image_size=(256, 256), softmax=True, trainable_softmax=False,
             final_conv_kernel=1, init_weights=False):
    if width not in (18, 32, 48):
        raise ValueError("width must be 18, 32 or 48")
    ch = [width, 2 * width, 4 * width, 8 * width]

    def stage(nmod, nb):
        return {"NUM_MODULES": nmod, "NUM_BRANCHES": nb, "BLOCK": "BASIC", "NUM_BLOCKS": [4] * nb,
                "NUM_CHANNELS": ch[:nb], "FUSE_METHOD": "SUM"}

    return to_cfg({
        "MODEL": {
            "NAME": "pose_hrnet_softmax" if softmax else "pose_hrnet",
            "INIT_WEIGHTS": init_weights, "PRETRAINED": "", "NUM_JOINTS": num_joints,
            "IMAGE_SIZE": list(image_size), "HEATMAP_SIZE": [image_size[0] // 4, image_size[1] // 4],
            "HEATMAP_SOFTMAX": softmax, "TRAINABLE_SOFTMAX": trainable_softmax, "SIGMA": 2,
            "EXTRA": {
                "PRETRAINED_LAYERS": ["conv1", "bn1", "conv2", "bn2", "layer1", "transition1", "stage2",
                                      "transition2", "stage3", "transition3", "stage4"],
                "FINAL_CONV_KERNEL": final_conv_kernel,
                "STAGE2": stage(1, 2), "STAGE3": stage(4, 3), "STAGE4": stage(3, 4),
            },
        },
        "LOSS": {"WITH_HEATMAP_LOSS": True, "HEATMAP_LOSS_FACTOR": 1.0, "WITH_POSE2D_LOSS": True,
                 "POSE2D_LOSS_FACTOR": 0.1},
        "TEST": {"POST_PROCESS": True},
    })


def load_yaml(path):
    import yaml
    with open(path) as f:
        cfg = to_cfg(yaml.safe_load(f))
    cfg.MODEL.setdefault("NUM_JOINTS", 21)
    return cfg
