"""Data-parallel plumbing for the path: images are independent units, so inference shards the batch across ranks
with NO data-path collective; only timing (max over ranks) and the optional gather of decoded joints use
torch.distributed.  Training adds the path's one real exchange step: the sum-all-reduce of the flat fp32 gradient
buffer (NCCL over NVLink on the GPU box), whose mean is taken by the optimizer kernel's grad_scale = 1/world.
(The reference's own multi-GPU runtime is nn.DataParallel / optional DDP, tools/train.py:221-254: DDP averages
gradients over ranks, BatchNorm statistics stay per rank - MODEL.SYNC_BN defaults to False, config/default.py:59.)"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one extra)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu", group=None):
    """All-reduce MAX of a python float (step time): the slowest rank defines the job's time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_joints(local, n_total, group=None):
    """All-gather per-rank decoded joints [b_r, J, 2] (ragged over ranks) into [n_total, J, 2] in shard order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


class GradAllReduce:
    """Sum-all-reduce of the flat gradient buffer.  `mean_scale` is what FlatParams.set_grad_scale must be given so that
    the optimizer sees the DDP-style mean.

    Two forms:
      * `ar(flat_grads)`: the whole buffer in `n_buckets` contiguous buckets on the current stream (host logic / CPU tests);
      * overlapped (CUDA, `device` given): `launch(flat_grads, lo, hi)` all-reduces one bucket on a communication stream
        that first waits for everything queued on the current stream so far, `finish()` makes the current stream wait for
        all launched buckets.  TrainEngine.train_step calls `launch` after each backward segment (head + stage 4, stage 3,
        the rest), so every bucket but the last small one travels under the remaining backward kernels."""

    def __init__(self, n_elems, n_buckets=1, group=None, device=None):
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        n_buckets = max(1, int(n_buckets))
        step = (n_elems + n_buckets - 1) // n_buckets
        step = (step + 127) // 128 * 128
        self.bounds = [(lo, min(n_elems, lo + step)) for lo in range(0, n_elems, step)]
        self.overlap = device is not None and torch.device(device).type == "cuda"
        self.comm = torch.cuda.Stream(device=device) if self.overlap else None
        self.device = device
        self.launched = []

    @property
    def mean_scale(self):
        return 1.0 / self.world

    def __call__(self, flat_grads):
        if self.world == 1:
            return flat_grads
        for lo, hi in self.bounds:
            dist.all_reduce(flat_grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        return flat_grads

    def launch(self, flat_grads, lo, hi):
        self.launched.append((lo, hi))
        if self.world == 1 or hi <= lo:
            return
        main = torch.cuda.current_stream(self.device)
        self.comm.wait_stream(main)
        with torch.cuda.stream(self.comm):
            dist.all_reduce(flat_grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group)

    def finish(self):
        self.launched = []
        if self.world > 1:
            torch.cuda.current_stream(self.device).wait_stream(self.comm)
