"""Multi-GPU plumbing: one process per GPU, posterior samples sharded across ranks.

The predictive path has exactly one exchange step: each rank reduces its own
samples to per-point float64 (mean, M2); the shards' moments are all-gathered over NCCL/NVLink
(16 bytes per point per rank) and Chan-merged on every rank by the bnn_moments_merge kernel, so
all ranks finish with identical mean / unbiased variance.  Raw float32 (sum, sum^2) is NOT used:
it loses ~13% of the variance when sigma << |mean|.
"""
import torch
import torch.distributed as dist


def shard_range(S, rank, world):
    """Contiguous, balanced [lo, hi) of S samples for `rank` (first S % world ranks get one more)."""
    base, extra = divmod(int(S), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_shards(part_mean, part_m2, count, group=None, merge_fn=None):
    """All-gather (count, mean, M2) of every rank and merge.  Returns (mean, var) float32.

    merge_fn(part_mean[R, ...], part_m2[R, ...], counts) -> (mean, var); defaults to the CUDA
    kernel (bnn_b200.ops.moments_merge).  There is no CPU implementation in this package; the
    gloo/CPU unit test injects the oracle's merge to exercise the sharding + collective logic.
    """
    if merge_fn is None:
        from . import ops
        merge_fn = ops.moments_merge
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_fn(part_mean.unsqueeze(0), part_m2.unsqueeze(0), [int(count)])
    R = dist.get_world_size(group)
    shape = tuple(part_mean.shape)
    packed = torch.cat([part_mean.reshape(-1), part_m2.reshape(-1)]).contiguous()      # [2*M] float64
    gathered = torch.empty(R * packed.numel(), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    gathered = gathered.view(R, 2, -1)
    cnt = torch.tensor([int(count)], dtype=torch.int64, device=packed.device)
    cnts = torch.empty(R, dtype=torch.int64, device=packed.device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    pm = gathered[:, 0].contiguous().view((R,) + shape)
    pm2 = gathered[:, 1].contiguous().view((R,) + shape)
    return merge_fn(pm, pm2, cnts.tolist())
