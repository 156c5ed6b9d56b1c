"""Worker of tests/test_gpu_multi.py (one process per GPU, NCCL): data-parallel training semantics of the reference's DDP path
(tools/train.py:239-244: per-rank BatchNorm statistics, gradients averaged over ranks, identical optimizer step everywhere).

Checks, on `world` GPUs:
  1. after the bucketed, overlapped NCCL all-reduce the flat gradient buffer holds the SUM of the per-rank gradients: rank 0
     recomputes every rank's shard gradient locally (same engine, no collective) and compares;
  2. after the fused Adam step (grad_scale = 1/world) all ranks hold BIT-IDENTICAL weights, and they equal the weights a
     single process gets from the mean of the shard gradients;
  3. the overlapped path gives the same result as the plain single-collective path.
Prints MULTI_GPU_OK on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from hrnet_b200 import synthetic
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet_softmax
    from hrnet_b200.parallel import GradAllReduce
    from hrnet_b200.train import TrainEngine
    B, H, W = 4, 128, 128

    def engine():
        torch.manual_seed(0)
        m = pose_hrnet_softmax.get_pose_net(make_cfg(32, trainable_softmax=True, image_size=(H, W)), is_train=False).to(dev).train()
        return m, TrainEngine(m, lr=1e-3, weight_decay=1e-4)

    def shard(r):
        gt, xy, vis = synthetic.targets(B, 21, H // 4, W // 4, seed=50 + r)
        return [t.to(dev) for t in (synthetic.images(B, H, W, seed=40 + r), gt, xy, vis)]

    # --- data-parallel step, overlapped buckets ---
    m, eng = engine()
    ar = GradAllReduce(eng.flat.grads.numel(), device=dev)
    eng.flat.set_grad_scale(ar.mean_scale)
    for it in range(2):            # second iteration replays the captured segment graphs
        m2, eng2 = (m, eng) if it == 0 else engine()
        if it == 1:
            ar = GradAllReduce(eng2.flat.grads.numel(), device=dev)
            eng2.flat.set_grad_scale(ar.mean_scale)
            eng2.train_step(*shard(rank), allreduce=ar, optimizer_step=False)      # capture
        p = eng2.train_step(*shard(rank), allreduce=ar, optimizer_step=False)
        torch.cuda.synchronize()
        summed = eng2.flat.grads.clone()
        segs = p.ar_segments()
        assert segs[0][3] == eng2.flat.n_grads and segs[-1][2] == 0 and all(a[2] == b[3] for a, b in zip(segs[:-1], segs[1:])), segs
        # --- every shard's gradient recomputed locally, no collective ---
        m3, eng3 = engine()
        local = torch.zeros_like(summed)
        for r in range(world):
            eng3.train_step(*shard(r), optimizer_step=False)
            local += eng3.flat.grads
        torch.cuda.synchronize()
        scale = float(local.abs().max())
        err = float((summed - local).abs().max()) / scale
        assert err < 1e-4, ("all-reduced gradient != sum of shard gradients", it, err)
    # --- plain single-collective path gives the same sum ---
    m4, eng4 = engine()
    eng4.train_step(*shard(rank), allreduce=GradAllReduce(eng4.flat.grads.numel()), optimizer_step=False)
    torch.cuda.synchronize()
    assert float((eng4.flat.grads - summed).abs().max()) / scale < 1e-4
    # --- optimizer step: identical weights on every rank ---
    m5, eng5 = engine()
    ar5 = GradAllReduce(eng5.flat.grads.numel(), device=dev)
    eng5.flat.set_grad_scale(ar5.mean_scale)
    for _ in range(3):
        eng5.train_step(*shard(rank), allreduce=ar5)
    torch.cuda.synchronize()
    w = eng5.flat.data.clone()
    ref = w.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(w, ref), "ranks diverged after the optimizer step"
    bufs = torch.cat([b.float().reshape(-1) for n, b in m5.named_buffers() if n.endswith("running_mean")])
    bref = bufs.clone()
    dist.broadcast(bref, src=0)
    if rank != 0:
        assert not torch.equal(bufs, bref), "BatchNorm statistics must stay per rank (SYNC_BN: false)"
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK world=%d grad_err=%.2e" % (world, err))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
