"""Data-parallel plumbing for the hot path: one process per GPU, rows sharded, weights
replicated.  H1 and H2 need no collective; H3 has one exchange step, the sum all-reduce of the flat gradient.
Everything here is backend-agnostic torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n_rows, world, rank):
    """Contiguous, balanced row range [lo, hi) of `rank` (the first n_rows % world ranks get one more row)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_segments(seg_off, world, rank):
    """Split every mode segment of a mode-sorted batch across ranks (H2): returns (row index list of this
    rank as [lo, hi) pairs per mode, local seg_off, global rows per mode for the reference's 1/B factor)."""
    pairs, local_off, counts = [], [0], []
    for m in range(len(seg_off) - 1):
        n = seg_off[m + 1] - seg_off[m]
        lo, hi = shard_bounds(n, world, rank)
        pairs.append((seg_off[m] + lo, seg_off[m] + hi))
        local_off.append(local_off[-1] + hi - lo)
        counts.append(n)
    return pairs, local_off, counts


def global_inv_count(local_rows, action_dim, group=None):
    """1 / (B_global * A): the mse_loss normaliser when every rank holds `local_rows` rows of one batch."""
    _, world = world_info(group)
    t = torch.tensor([float(local_rows)], dtype=torch.float64)
    if world > 1:
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, group=group)
    return 1.0 / (t.item() * action_dim)


def allreduce_sum_(flat_grads, loss=None, group=None):
    """The single exchange step of H3: in-place sum of the flat fp32 gradient (and the loss partial sum)."""
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(flat_grads, group=group)
        if loss is not None:
            dist.all_reduce(loss, group=group)
    return flat_grads, loss
