"""Multi-GPU plumbing of the box path: images are independent units, sharded by rank; the only
cross-image coupling of MultiBoxLoss is the normaliser N = sum_b num_pos and the two loss
numerators (multibox_loss.py:114-116), so ONE all-reduce(SUM) of three fp64 scalars per step is
the whole data-path communication.  Detect needs none."""
import torch


def shard_range(num_images, rank, world):
    """[begin, end) of the images rank `rank` owns (contiguous, sizes differ by at most one)."""
    base, rem = divmod(int(num_images), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(loc, conf, targets, rank, world):
    b, e = shard_range(loc.size(0), rank, world)
    return loc[b:e], conf[b:e], targets[b:e]


def allreduce_loss_sums(sums, group=None):
    """sums: fp64 [3] = {sum smooth-L1, sum CE, N_pos} of the local shard -> global sums (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def finalize_losses(sums):
    """(loss_l, loss_c) = sums[0:2] / N with the N == 0 convention of the library (zeros)."""
    n = float(sums[2])
    if n <= 0:
        return torch.zeros((), dtype=torch.float32), torch.zeros((), dtype=torch.float32)
    return (sums[0] / n).to(torch.float32), (sums[1] / n).to(torch.float32)
