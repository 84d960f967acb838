"""Multi-GPU layout of the hot path: scenes are independent units, so ranks own disjoint scene
ranges and no data-path collective exists.  The only cross-rank traffic of
the op benchmarks is the max-over-ranks reduction of the timed region."""
from typing import Tuple

import torch
import torch.distributed as dist


def scene_range(rank: int, world: int, scenes_per_rank: int) -> Tuple[int, int]:
    """Weak scaling: rank r works on scene ids [r*S, (r+1)*S)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return rank * scenes_per_rank, (rank + 1) * scenes_per_rank


def split_scenes(num_scenes: int, world: int):
    """Strong scaling: contiguous, balanced split of a fixed scene list (first ranks get the remainder)."""
    base, rem = divmod(num_scenes, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((start, start + n))
        start += n
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed milliseconds -> the slowest rank's figure (what the whole job waits for)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank: int, world: int, ms: float) -> float:
    """Whole-job units per second for a weak-scaled step that took `ms` on the slowest rank."""
    return units_per_rank * world / (ms / 1e3)
