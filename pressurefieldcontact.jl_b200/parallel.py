"""Multi-GPU plumbing (one process per GPU, torch.distributed): the two ways the path shards.

* Batched independent scenarios (MPC roll-outs): contiguous environment ranges per rank, no
  collective on the hot path (`env_range`, `gather_env_results`).
* One very large scene: every rank traverses and sorts (identical pair lists), evaluates its slice
  of the 256-pair chunks, and the per-instruction partial sums are all-reduced (`eval_sharded`):
  one exchange for regularized friction, three for bristle.

The reference has no parallelism of any kind; these are the natural shardings of
forceAllElasticIntersections! (src/contact_algorithms_non_friction.jl:60-68).
"""
from __future__ import annotations

import numpy as np


def env_range(n_env: int, rank: int, world: int):
    """Contiguous, balanced environment range [lo, hi) of `rank`."""
    lo = (n_env * rank) // world
    hi = (n_env * (rank + 1)) // world
    return lo, hi


def gather_env_results(local: "np.ndarray", n_env: int, group=None):
    """all_gather of per-environment results (first axis = this rank's environments) into the full batch."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [env_range(n_env, r, world)[1] - env_range(n_env, r, world)[0] for r in range(world)]
    t = torch.as_tensor(np.ascontiguousarray(local))
    pad = max(sizes)
    if t.shape[0] < pad:  # all_gather wants equal shapes: pad the short ranks, trim after
        t = torch.cat([t, torch.zeros((pad - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype)], dim=0)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return torch.cat([o[:sizes[r]] for r, o in enumerate(out)], dim=0).numpy()


def allreduce_sum_(tensor, group=None):
    """In-place sum over ranks of the (small) partial-sum buffer."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def eval_sharded(ctx, n_env, X, twist, s, wrench, sdot, n_pairs, flags, reduce_partials):
    """Drives the sharded protocol of include/pfc.h on device buffers (torch CUDA tensors or None).
    reduce_partials(ptr, count) must sum `count` doubles at device address `ptr` over all ranks, ordered
    after the work already queued on ctx.stream (e.g. a torch.distributed all_reduce issued on that stream)."""
    p = lambda t: None if t is None else t.data_ptr()
    ctx.eval_sharded_begin(n_env, p(X), p(twist), p(s), p(wrench), p(sdot), p(n_pairs), p(flags))
    n_exchange = 0
    while True:
        ptr, count = ctx.eval_sharded_partials()
        if count:
            reduce_partials(ptr, count)
            n_exchange += 1
        if not ctx.eval_sharded_step():
            break
    return n_exchange
