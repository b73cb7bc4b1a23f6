"""Session sharding across the GPUs of one box.

Sessions are independent units -- a session's KV ring, adapter cache, fbank carry and pe_index are touched by that
session only (bin/dialog_state_pred.py:221-232, 793-814 in the reference) -- so the multi-GPU design is a static
partition of sessions over ranks with replicated weights and NO collective on the data path.  The only exchange is the
gather of per-rank counters at the end of a run (torch.distributed; NCCL on GPUs, gloo in the CPU tests).
The reference's own placement policy is "least loaded pipeline" (bin/pool.py:79-83); `assign_least_loaded` mirrors it
for sessions that arrive over time, `partition` is the static round-robin used by the benchmark.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import torch


def partition(session_ids: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Static round-robin partition: session i of the sorted id list goes to rank i % world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return [s for i, s in enumerate(sorted(session_ids)) if i % world_size == rank]


def owner(session_ids: Sequence[int], world_size: int) -> Dict[int, int]:
    return {s: i % world_size for i, s in enumerate(sorted(session_ids))}


def assign_least_loaded(loads: Sequence[int]) -> int:
    """Rank that should take the next arriving session (bin/pool.py:79-83: the pipeline with the fewest users)."""
    best, arg = None, 0
    for r, n in enumerate(loads):
        if best is None or n < best:
            best, arg = n, r
    return arg


def gather_stats(local: Dict[str, float], group=None) -> Dict[str, float]:
    """End-of-run gather: sums of additive counters, max of every key that starts with 'max_' (timings are reported
    as the max over ranks).  One small all_reduce pair; not on the hot path."""
    import torch.distributed as dist
    keys = sorted(local)
    add = torch.tensor([float(local[k]) for k in keys if not k.startswith("max_")], dtype=torch.float64)
    mx = torch.tensor([float(local[k]) for k in keys if k.startswith("max_")], dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        add, mx = add.to(dev), mx.to(dev)
        if add.numel():
            dist.all_reduce(add, op=dist.ReduceOp.SUM, group=group)
        if mx.numel():
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        add, mx = add.cpu(), mx.cpu()
    out, ia, im = {}, 0, 0
    for k in keys:
        if k.startswith("max_"):
            out[k] = float(mx[im]); im += 1
        else:
            out[k] = float(add[ia]); ia += 1
    return out


def throughput(stats: Dict[str, float]) -> float:
    """Whole-job audio-seconds per second: all ranks' audio divided by the slowest rank's time."""
    return stats["audio_seconds"] / stats["max_elapsed_s"] if stats.get("max_elapsed_s") else 0.0
