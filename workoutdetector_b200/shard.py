"""Multi-GPU sharding of the embarrassingly parallel path: one process per GPU, each with its own
engine and replicated weights; the unit of work is a video (its windows share decoded frames) or a clip. There is no
data-path collective — per-video results travel to rank 0 over a host-side (gloo) object gather.
The reference has no multi-GPU inference at all (inference_dataset is a serial loop, utils/inference_count.py:399)."""
from typing import Dict, List, Optional, Sequence

import torch.distributed as dist


def partition_lpt(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first greedy: item indices per rank, balanced by cost (e.g. total_frames).
    Deterministic: ties go to the lower rank, items inside a shard keep ascending index order."""
    loads = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        r = min(range(world), key=lambda r: (loads[r], r))
        shards[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in shards]


def partition_contiguous(n: int, world: int) -> List[range]:
    """n equal-cost items in contiguous blocks whose sizes differ by at most one."""
    base, extra = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append(range(start, start + size))
        start += size
    return out


_host_group = None


def host_group():
    """A gloo group for host-side object gathers (created once; the default group may be NCCL)."""
    global _host_group
    if _host_group is None:
        _host_group = dist.new_group(backend="gloo") if dist.get_backend() != "gloo" else dist.group.WORLD
    return _host_group


def gather_to_rank0(local: Dict) -> Optional[Dict]:
    """Merge every rank's {video_name: result} dict on rank 0 (None elsewhere). Single-process: returns local."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    grp = host_group()
    buf = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, buf, dst=0, group=grp)
    if dist.get_rank() != 0:
        return None
    merged: Dict = {}
    for part in buf:
        overlap = merged.keys() & part.keys()
        if overlap:
            raise RuntimeError(f"videos assigned to more than one rank: {sorted(overlap)[:5]}")
        merged.update(part)
    return merged
