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
