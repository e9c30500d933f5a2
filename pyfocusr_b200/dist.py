"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for rendezvous and timing.

The spectral stage shards over independent mesh pairs: pair ``i`` goes to
rank ``i mod world``; no collective touches the data path.  The only collectives are the ones the
measurement needs (barrier, max-over-ranks of the device time, sum of counters) and an optional
gather of the small per-pair results.  The row-partitioned >=1M-vertex solve (section 8e-ii) is the
only part of the path with a real exchange step; DESIGN.md lists it under "next".

Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import os

__all__ = ["pair_shard", "sub_batches", "init", "world", "barrier", "all_reduce_max", "all_reduce_sum", "gather_objects", "finalize"]


def pair_shard(n_pairs_total, rank, world_size):
    """Pair ids owned by ``rank``: i with i mod world_size == rank."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, n_pairs_total, world_size))


def sub_batches(pair_ids, n_sub):
    """A rank's pair ids as ``n_sub`` contiguous, non-empty, near-equal groups (fewer if there are fewer pairs): the
    sub-batches ``SpectralBatch.run_concurrent`` keeps in flight on one GPU."""
    ids = list(pair_ids)
    if not ids:
        return []   # a rank without pairs has no sub-batch to run (callers skip the step)
    n_sub = max(1, min(int(n_sub), len(ids)))
    return [ids[k * len(ids) // n_sub:(k + 1) * len(ids) // n_sub] for k in range(n_sub)]


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process if absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend=None):
    """Initialise the default process group when launched under torchrun with WORLD_SIZE > 1."""
    import torch
    import torch.distributed as dist

    rank, local, size = world()
    if size <= 1 or dist.is_initialized():
        return rank, local, size
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group(backend, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return rank, local, size


def _active():
    import torch.distributed as dist

    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def barrier():
    import torch

    if _active():
        import torch.distributed as dist

        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def _reduce(value, op_name):
    import torch

    if not _active():
        return float(value)
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def all_reduce_max(value):
    """Max over ranks (every multi-GPU time is reported as the slowest rank's device time)."""
    return _reduce(value, "MAX")


def all_reduce_sum(value):
    return _reduce(value, "SUM")


def gather_objects(obj):
    """List of every rank's ``obj`` on every rank (small per-pair results only)."""
    if not _active():
        return [obj]
    import torch.distributed as dist

    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def finalize():
    if _active():
        import torch.distributed as dist

        dist.destroy_process_group()
