"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  Replaces the reference's single-process ``torch.nn.DataParallel`` wrap (``executor/executor.py:53-56``),
which re-broadcasts all 1.84 M parameters on every forward and reduces gradients into GPU 0 on every backward:
here weights are replicated once, point batches are sharded by contiguous rows, and one fp32 sum-allreduce of the flat
gradient vector (+ one 8-double message of loss sums) per step is the only exchange.  Grid and octree inference
shard by slab / subtree with no collective at all.
"""
import torch
import torch.distributed as dist


def is_on():
    return dist.is_available() and dist.is_initialized()


def world_size():
    return dist.get_world_size() if is_on() else 1


def rank():
    return dist.get_rank() if is_on() else 0


def all_reduce_sum(t):
    """In-place sum over ranks (no-op on a single process)."""
    if is_on() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_rows(n, rank_=None, world=None):
    """Contiguous row range [lo, hi) of an n-row batch owned by ``rank_`` (contiguous N/G rows)."""
    world = world_size() if world is None else world
    rank_ = rank() if rank_ is None else rank_
    base, rem = divmod(int(n), int(world))
    lo = rank_ * base + min(rank_, rem)
    return lo, lo + base + (1 if rank_ < rem else 0)


def slab_bounds(n_axis, rank_=None, world=None):
    """Planes [ix0, ix1) of the slowest grid axis owned by ``rank_`` (reference executor/executor.py:351-356 lays the
    flattened grid out with x slowest, so the contiguous slab BASELINE.json calls a "z-slab" is a range of ix)."""
    return shard_rows(n_axis, rank_, world)


def subtree_owner(root_index, world=None):
    """Octree subtrees (cells of the uniform level) are dealt round-robin to ranks."""
    world = world_size() if world is None else world
    return int(root_index) % int(world)


def broadcast_parameters(module, src=0):
    """Replicate the weights once (instead of DataParallel's broadcast per forward)."""
    if is_on() and dist.get_world_size() > 1:
        with torch.no_grad():
            for p in module.parameters():
                # p.detach() shares p's version counter (p.data does not): the in-place receive bumps it, so the engine's
                # packed-weight cache sees the new values
                dist.broadcast(p.detach(), src=src)
    inner = module.module if hasattr(module, "module") else module
    if hasattr(inner, "invalidate_packed"):
        inner.invalidate_packed()
