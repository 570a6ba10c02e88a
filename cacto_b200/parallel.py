"""Data-parallel plumbing of the hot path (one process per GPU, torch.distributed).

The path shards two ways:
  * rollouts / Jacobians / reward-to-go are independent units -> contiguous block partition of the unit index,
    replicated actor weights, NO collective on the data path (``shard_range``);
  * the update is synchronous data parallelism: every rank takes B/world rows of the same globally sampled
    minibatch, computes gradient SUMS scaled by 1/B_global, and one all-reduce(sum) per network gives every
    replica the full-batch gradient before the identical Adam step (``allreduce_sum``).  Critic and actor
    steps are sequentially dependent (RL.py:104-109), hence one collective per network.
The functions take the process-group module/object so that the same code runs over NCCL on GPUs and over
gloo in the CPU tests.
"""


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of ``n`` units owned by ``rank``; the first n % world ranks get one extra."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum(tensor, dist):
    """In-place sum over the group (NCCL over NVLink on GPUs).  No-op without a group or with one rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return tensor
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def global_batch(local_batch, dist):
    if dist is None or not dist.is_initialized():
        return int(local_batch)
    return int(local_batch) * dist.get_world_size()
