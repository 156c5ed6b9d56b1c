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
    """Sum-all-reduce of the flat gradient buffer in `n_buckets` contiguous buckets (bucket boundaries aligned to 128
    elements).  With one bucket this is a single collective per step; more buckets let the caller overlap the exchange
    of already finished buckets with the rest of the backward pass.  `mean_scale` is what FlatParams.set_grad_scale
    must be given so that the optimizer sees the DDP-style mean."""

    def __init__(self, n_elems, n_buckets=1, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        n_buckets = max(1, int(n_buckets))
        step = (n_elems + n_buckets - 1) // n_buckets
        step = (step + 127) // 128 * 128
        self.bounds = [(lo, min(n_elems, lo + step)) for lo in range(0, n_elems, step)]

    @property
    def mean_scale(self):
        return 1.0 / self.world

    def __call__(self, flat_grads):
        if self.world == 1:
            return flat_grads
        for lo, hi in self.bounds:
            dist.all_reduce(flat_grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        return flat_grads
