"""Multi-GPU plumbing: sessions are independent, so a batch is block-partitioned across one process per GPU with no
data-plane collective (the reference has no collectives at all; its only scale-out is Ray actor replicas,
/root/reference/Cluster/InfernTTSActor.py:12).  The single collective is the control-plane gather of per-GPU stats."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

STAT_FIELDS = ("sessions", "steps", "g711_bytes", "kernel_launches", "device_ms")


def shard_range(n_sessions: int, world: int, rank: int) -> Tuple[int, int]:
    """Block partition: rank r owns sessions [start, start+count).  Counts differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_sessions, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def session_rank(session_index: int, n_sessions: int, world: int) -> int:
    """Inverse of shard_range: the rank that owns a session (a session never migrates: its pre_frames slot lives in
    one GPU's HBM)."""
    base, rem = divmod(n_sessions, world)
    edge = rem * (base + 1)
    if session_index < edge:
        return session_index // (base + 1)
    return rem + (session_index - edge) // max(base, 1)


def gather_stats(local: Dict[str, float], device=None) -> List[Dict[str, float]]:
    """all_gather of a fixed-size stats struct; NCCL on GPUs, gloo in the CPU tests.  Returns one dict per rank."""
    vec = torch.tensor([float(local.get(k, 0.0)) for k in STAT_FIELDS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.zeros_like(vec) for _ in range(dist.get_world_size())]
        dist.all_gather(out, vec)
    else:
        out = [vec]
    return [{k: float(v[i]) for i, k in enumerate(STAT_FIELDS)} for v in out]


def max_over_ranks(x: float, device=None) -> float:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x
