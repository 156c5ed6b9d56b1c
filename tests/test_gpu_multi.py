"""-m gpu, needs >= 2 devices: NCCL data-parallel equivalence on hardware (SURVEY §4 tier 4; the reference's DDP path,
tools/train.py:239-244).  The checks live in tests/multi_gpu_worker.py (one process per GPU, launched with torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 CUDA devices")
def test_two_rank_nccl_gradient_and_weight_equivalence():
    env = dict(os.environ, NCCL_DEBUG="WARN", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 CUDA devices")
def test_nn_data_parallel_eval_forward_on_two_devices():
    """model.eval() under nn.DataParallel over two GPUs == the single-GPU forward (per-device engines, one host thread per GPU)"""
    import sys
    sys.path.insert(0, ROOT)
    from hrnet_b200 import synthetic
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet_softmax
    torch.manual_seed(0)
    m = pose_hrnet_softmax.get_pose_net(make_cfg(32, image_size=(128, 128)), is_train=False).cuda(0).eval()
    x = synthetic.images(4, 128, 128).cuda(0)
    ref = m(x)[0].clone()
    dp = torch.nn.DataParallel(m, device_ids=[0, 1])
    for _ in range(2):
        out = dp(x)[0]
        assert out.shape == ref.shape
        # each replica sees a batch of 2 (tile shapes may differ from the batch-4 plan): bf16 noise, not bit equality
        assert (out - ref).abs().max().item() <= 5e-2 * ref.abs().max().item()
    assert len(m._shared["engines"]) == 2
