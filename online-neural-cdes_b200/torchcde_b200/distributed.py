"""Data-parallel use of the solve: shard the batch of series over ranks, one NCCL all-reduce of the gradients.

The reference has no multi-GPU path inside a run (GNU-parallel of independent processes).  The batch
elements of a fixed-grid solve are independent, so each rank solves its shard with no data-path collective; the only
exchange is the sum of the parameter gradients after backward.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world_size):
    """Contiguous shard [lo, hi) of n series for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank=None, world_size=None):
    """Slice every tensor in `tensors` along dim 0 to this rank's shard."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    out = []
    for t in tensors:
        lo, hi = shard_bounds(t.shape[0], rank, world_size)
        out.append(t[lo:hi])
    return out


def allreduce_gradients(params, average=False, group=None):
    """ONE collective over a flat buffer holding every gradient (sum, or mean over ranks when average=True).

    Parameters without a gradient contribute zeros so that all ranks issue the same collective."""
    params = [p for p in params if p.requires_grad]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel() * flat.element_size()
