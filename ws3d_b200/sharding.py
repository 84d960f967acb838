"""Multi-GPU layout of the hot path: scenes are independent units, so ranks own disjoint scene
ranges and no data-path collective exists.  The only cross-rank traffic of
the op benchmarks is the max-over-ranks reduction of the timed region.  Training (BASELINE config 3) adds the path's one
collective, the data-parallel gradient all-reduce: `FlatGradients`."""
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


class FlatGradients:
    """Every parameter gradient of a replica as a view into ONE contiguous FP32 buffer, averaged over the ranks with a
    single all-reduce per step (NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU tests).

    This is the exchange the reference gets from its data-parallel wrapper (tools/train_rpn.py wraps the model in
    nn.DataParallel; the DDP form of the same thing is one averaged all-reduce of the gradients), restated so that it
    fits in a CUDA graph: no bucket hooks, no per-step Python, one collective of `nbytes` (12.2 MB for the Stage-1 RPN,
    far below the size where NVLink time matters, so it is issued once after backward instead of being overlapped bucket
    by bucket).  autograd accumulates into the views in place (AccumulateGrad adds into a defined .grad), `zero()` is one
    memset, the optimiser reads the same views."""

    def __init__(self, params, world: int = None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        if any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise ValueError("FlatGradients wants float32 parameters on one device")
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        live = dist.is_available() and dist.is_initialized()
        self.world = (dist.get_world_size() if live else 1) if world is None else int(world)
        if self.world > 1 and not live:
            raise RuntimeError("world > 1 without an initialised process group")
        self._avg = self.world > 1 and dist.get_backend() == "nccl"

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self) -> None:
        self.flat.zero_()

    def attached(self) -> bool:
        """True while every .grad still is this buffer's view (an optimiser's zero_grad(set_to_none=True) detaches them)."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                return False
            off += p.numel()
        return True

    def exchange(self) -> None:
        """Mean over the ranks, in place.  One collective; a no-op on one rank."""
        if self.world == 1:
            return
        if self._avg:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(self.world)
