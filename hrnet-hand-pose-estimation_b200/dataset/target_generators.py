"""Ground-truth heat maps on the device: drop-in for lib/dataset/target_generators/target_generators.py:15-53.

    HeatmapGenerator(output_res, num_joints, sigma=-1)(joints)

The reference builds the maps per sample with numpy inside the DataLoader workers and ships 21 x 64 x 64 fp32 per image to the
GPU (22 MB of the 72 MB a batch-64 training step uploads).  Here `joints` is a CUDA tensor [J, 3] or [B, J, 3] of (u, v, visible)
rows in heat-map pixels and the maps are written by one kernel launch (hrnb_gen_heatmaps): the host only uploads B*J*3 floats.
Semantics are the reference's, bit for bit up to float32 rounding of exp: x = int(u), y = int(v); invisible (visible <= 0) or
out-of-map joints give a zero map; a (6 sigma + 3)^2 Gaussian patch, peak 1, everything else zero.  `output_res` may be an int
(square, as in the reference) or (h, w).  CUDA tensors only - no CPU fallback."""
import torch

from .. import _lib


class HeatmapGenerator:
    def __init__(self, output_res, num_joints, sigma=-1):
        self.h, self.w = (output_res, output_res) if isinstance(output_res, int) else tuple(output_res)
        self.output_res = output_res
        self.num_joints = num_joints
        if sigma < 0:
            sigma = self.h / 64
        self.sigma = float(sigma)

    def __call__(self, joints, out=None):
        if not torch.is_tensor(joints) or not joints.is_cuda:
            raise RuntimeError("the B200 HeatmapGenerator takes CUDA tensors (no CPU fallback); joints: [J, 3] or [B, J, 3]")
        single = joints.dim() == 2
        j = joints.unsqueeze(0) if single else joints
        B, J, c = j.shape
        assert J == self.num_joints and c in (2, 3), (tuple(joints.shape), self.num_joints)
        j = j.contiguous().float()
        if out is None:
            out = torch.empty((B, J, self.h, self.w), dtype=torch.float32, device=j.device)
        assert out.is_contiguous() and tuple(out.shape) == (B, J, self.h, self.w)
        with torch.cuda.device(j.device):
            _lib.check(_lib.lib().hrnb_gen_heatmaps(j.data_ptr(), c, B * J, self.h, self.w, self.sigma, out.data_ptr(),
                                                    _lib.stream_ptr()))
        return out[0] if single else out
