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
