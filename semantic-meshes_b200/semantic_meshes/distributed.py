"""View-sharded multi-GPU fusion (one process per GPU, torch.distributed).

The reference is single-process / single-GPU. Every view is independent for render, for the per-view pixel counts and for
the weights (include/semantic_meshes/fusion/Mesh.h:90-104 uses counts of THIS view only), and every aggregator kind
accumulates by addition, so views can be dealt to ranks in any way; each rank owns a private accumulator and ONE
all-reduce (sum) at the end makes every rank hold the accumulator of all views. No other communication exists.
"""


def shard_views(n_views, rank, world_size):
    """Indices of the views rank `rank` processes: round-robin, so ranks stay within one view of each other and
    consecutive (similar) camera poses spread over the GPUs."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_views, world_size))


def allreduce_accumulator(acc, group=None):
    """Sum a raw accumulator tensor over all ranks in place (NCCL on GPU tensors, gloo on CPU tensors in the tests).
    `mul` accumulators hold -log p with +inf as the absorbing zero; inf + x = inf keeps that through the sum."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def row_slices(n_rows, world_size):
    """Equal row blocks (the last ones padded): rank r owns rows [r * per, min((r + 1) * per, n_rows))."""
    per = (n_rows + world_size - 1) // world_size
    return per, [(min(r * per, n_rows), min((r + 1) * per, n_rows)) for r in range(world_size)]


def reduce_scatter_rows(acc, group=None):
    """Sum the (P, Cpad) accumulators of all ranks and leave every rank with the rows of its own slice only: one
    reduce-scatter, (N-1)/N x P x Cpad floats on the wire per rank instead of the all-reduce's 2 (N-1)/N.
    -> ((first, last), tensor of those summed rows). Without a process group: the whole accumulator."""
    import torch
    import torch.distributed as dist
    P = acc.shape[0]
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return (0, P), acc
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per, slices = row_slices(P, world)
    src = acc
    if per * world != P:  # pad to equal blocks
        src = torch.zeros((per * world,) + tuple(acc.shape[1:]), dtype=acc.dtype, device=acc.device)
        src[:P] = acc
    first, last = slices[rank]
    if not acc.is_cuda:  # gloo (the CPU tests) has no reduce-scatter: all-reduce a copy and cut the slice out
        total = src.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        return (first, last), total[first:last]
    out = torch.empty((per,) + tuple(acc.shape[1:]), dtype=acc.dtype, device=acc.device)
    dist.reduce_scatter_tensor(out, src.contiguous(), op=dist.ReduceOp.SUM, group=group)
    return (first, last), out[:last - first]
