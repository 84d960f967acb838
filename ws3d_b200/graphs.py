"""CUDA-graph replay of a fixed-shape forward pass.

The backbone forward is ~100 dependent launches on two streams; replaying it as one CUDA graph removes
the per-launch host work and most of the inter-kernel gaps.  The graph owns static input / output
buffers (tensor maps and kernel arguments are baked into it, so addresses must not change), which is why
this is an explicit wrapper and not something the modules do behind the caller's back.

    runner = CudaGraphRunner(model, example_batch)      # eval / no-grad inference
    feats = runner(batch)                               # copies into the static input, replays
"""
from typing import Callable

import torch


class CudaGraphRunner:
    def __init__(self, fn: Callable, example: torch.Tensor, warmup: int = 3):
        assert example.is_cuda, "CUDA graphs need a CUDA tensor"
        self.fn = fn
        self.static_in = example.clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):   # lazily built state (folded weights, cached scratch, streams) must exist before capture
                fn(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        torch.cuda.synchronize(example.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = fn(self.static_in)

    def __call__(self, x: torch.Tensor):
        if x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
