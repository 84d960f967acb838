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


class PipelinedBackboneRunner:
    """Software-pipelined, graph-replayed forward for a stream of equally shaped batches.

    Level-1 furthest point sampling is a chain of npoint-1 dependent iterations: at the Stage-1 shapes it takes
    half of a forward pass while using 64 of the 148 SMs and no memory bandwidth, and everything else waits for
    it.  It depends on coordinates only, so one replayed graph does, side by side,

        high-priority stream :  level-1 FPS of batch i+1                      (latency-bound, 4 SMs per cloud)
        main (+ side) stream :  the rest of the forward pass of batch i       (bandwidth-bound)

    and a batch costs max(FPS, rest) instead of FPS + rest.  Two graphs alternate between two buffer sets.

        runner = PipelinedBackboneRunner(backbone, example)      # eval / no-grad inference
        runner.prefetch(batch0)                                  # pipeline fill: sampling of the first batch
        out0 = runner.step(batch1)                               # features of batch0; samples batch1 meanwhile
        out1 = runner.step(batch2) ...                           # a returned tensor is valid until the next-but-one step

    `fn(pointcloud, first_samples)` is the consumer (default: backbone.forward -> per-point features); `backbone`
    supplies sample_first_level.  Results are bit-identical to the unpipelined forward (same kernels, same inputs).
    Batches may come from pinned host memory: the copy of batch i+2 then runs on a copy stream during step i.
    """

    def __init__(self, backbone, example: torch.Tensor, fn: Callable = None, warmup: int = 2):
        assert example.is_cuda
        dev = example.device
        self.device = dev
        self.backbone = backbone
        self.fn = fn or (lambda pc, first: backbone(pc, first_samples=first)[1])
        self.stage = [example.clone(), example.clone()]     # where the caller's batches land (H2D / D2D target)
        self.inputs = [example.clone(), example.clone()]    # what the graphs read
        self.fps_stream = torch.cuda.Stream(device=dev, priority=-1)
        self.copy_stream = torch.cuda.Stream(device=dev)
        with torch.no_grad():
            self.samples = [backbone.sample_first_level(x) for x in self.inputs]
        cur = torch.cuda.current_stream(dev)
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm), torch.no_grad():
            for _ in range(warmup):
                self.fn(self.inputs[0], self.samples[0])
        cur.wait_stream(warm)
        torch.cuda.synchronize(dev)
        self.graphs, self.outputs = [], []
        for p in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g):
                cap = torch.cuda.current_stream(dev)
                fork = torch.cuda.Event()
                fork.record(cap)
                with torch.cuda.stream(self.fps_stream):
                    self.fps_stream.wait_event(fork)
                    self.inputs[1 - p].copy_(self.stage[1 - p])
                    self.samples[1 - p].copy_(backbone.sample_first_level(self.inputs[1 - p]))
                    join = torch.cuda.Event()
                    join.record(self.fps_stream)
                out = self.fn(self.inputs[p], self.samples[p])
                cap.wait_event(join)
            self.graphs.append(g)
            self.outputs.append(out)
        self.cur = 0          # buffer set whose samples are ready = the batch the next step() completes
        self.primed = False
        self._staged = [None, None]   # copy-stream events: stage[k] holds the caller's data
        self._done = [None, None]     # main-stream events: the replay that last read stage[k] has finished

    def _stage(self, k: int, batch: torch.Tensor):
        main = torch.cuda.current_stream(self.device)
        if batch.data_ptr() == self.stage[k].data_ptr():
            return
        with torch.cuda.stream(self.copy_stream):
            if self._done[k] is not None:
                self.copy_stream.wait_event(self._done[k])
            else:
                self.copy_stream.wait_stream(main)
            self.stage[k].copy_(batch, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._staged[k] = ev

    def prefetch(self, batch: torch.Tensor):
        """Pipeline fill: stage the first batch and sample it (eager, one FPS launch)."""
        main = torch.cuda.current_stream(self.device)
        k = self.cur
        self._stage(k, batch)
        if self._staged[k] is not None:
            main.wait_event(self._staged[k])
            self._staged[k] = None
        with torch.no_grad():
            self.inputs[k].copy_(self.stage[k])
            self.samples[k].copy_(self.backbone.sample_first_level(self.inputs[k]))
        self.primed = True

    def stage_next(self, batch: torch.Tensor):
        """Start copying the batch that the NEXT step() will sample (it may be called before that step, while the
        previous one is still running: the copy waits only for the replay that last read the staging buffer)."""
        self._stage(1 - self.cur, batch)

    def step(self, next_batch: torch.Tensor = None):
        """Completes the batch staged one call earlier; samples `next_batch` (or whatever stage_next() staged,
        or the previous contents of the staging buffer) for the next call."""
        assert self.primed, "call prefetch(first_batch) before the first step()"
        main = torch.cuda.current_stream(self.device)
        p = self.cur
        if next_batch is not None:
            self._stage(1 - p, next_batch)
        if self._staged[1 - p] is not None:
            main.wait_event(self._staged[1 - p])
            self._staged[1 - p] = None
        self.graphs[p].replay()
        ev = torch.cuda.Event()
        ev.record(main)
        self._done[1 - p] = ev     # this replay read stage[1-p]
        self.cur = 1 - p
        return self.outputs[p]


def _plan_tensors(plan):
    """The tensors of a coordinate-phase plan (nested dict / list / tuple) in a fixed order."""
    if isinstance(plan, torch.Tensor):
        return [plan]
    if isinstance(plan, dict):
        return [t for key in sorted(plan) for t in _plan_tensors(plan[key])]
    if isinstance(plan, (list, tuple)):
        return [t for item in plan for t in _plan_tensors(item)]
    return []


class StreamedBackboneRunner:
    """Deeper software pipeline for throughput: the COORDINATE phase of the forward pass (four FPS levels, ball
    queries, interpolation stencils -- backbone.coordinate_phase) runs `lookahead` batches ahead of the FEATURE phase
    (grouping + MLPs + interpolation -- backbone.feature_phase), each as its own CUDA graph on its own stream.

    PipelinedBackboneRunner is bounded by the level-1 FPS latency (2.1 ms at the Stage-1 shapes, on 64 SMs).  Here
    the samplers run in their throughput mode (native.set_fps_mode(1): the spatially bucketed kernel, ONE SM per
    cloud, 1.7x the latency at a quarter of the SM time) and `lookahead` coordinate phases are in flight at any time
    on a few SMs each, while the remaining SMs run the bandwidth-bound feature phase of the batch whose coordinates
    are ready.  A batch then costs max(coordinate phase / lookahead, feature phase) of wall time; its own latency
    grows to the sum of the two.

        runner = StreamedBackboneRunner(backbone, example)
        for b in the first `lookahead` batches: runner.submit(b)     # host (pinned) or device tensors
        loop:  out = runner.complete(); runner.submit(next_batch)    # `out` is valid until `lookahead` more completes

    Cold start: a batch submitted while fewer than `cold_start` batches are in flight has its coordinate phase on the critical
    path and most of the machine to itself, so it replays a second capture of the same phase taken with the LATENCY samplers
    (fps mode 0: 2.1 ms on 4 SMs per cloud instead of 3.2 ms on half an SM) whose results are copied into the same plan
    tensors.  Same indices either way (FPS is exact in every mode); `cold_start=0` disables it.

    `fn(pointcloud, plan)` is the consumer of the feature phase (default: backbone.feature_phase -> per-point
    features).  Results are bit-identical to the plain forward: FPS is exact in every mode and the rest is the same
    kernels on the same inputs.
    """

    def __init__(self, backbone, example: torch.Tensor, fn: Callable = None, lookahead: int = 4, fps_mode: int = 1,
                 feature_streams: int = 1, warmup: int = 2, cold_start: int = 2):
        from . import native
        assert example.is_cuda and lookahead >= 1
        # every buffer set captures its cell grids in its own scratch arena (arena 0 stays with eager callers): two
        # sets sharing an arena would replay on different coordinate streams over the same scratch addresses
        arenas = native.num_arenas()
        if lookahead + 1 > arenas - 1:
            raise ValueError(f"lookahead {lookahead} needs {lookahead + 1} scratch arenas; the library has {arenas - 1} besides "
                             "the default one (ws3d_num_arenas)")
        dev = example.device
        self.device, self.backbone, self.lookahead = dev, backbone, lookahead
        # feature_streams > 1: the feature phases of consecutive batches alternate between that many internal streams
        # (the deeper levels are chains of sub-wave kernels: two batches side by side fill the SMs one leaves idle);
        # complete(consume=...) then runs the caller's post-processing on the same stream and join() orders the
        # caller's stream after all of them.  1: the feature phase runs on the caller's current stream.
        self.feature_streams = ([torch.cuda.Stream(device=dev) for _ in range(feature_streams)] if feature_streams > 1 else None)
        self.fn = fn or (lambda pc, plan: backbone.feature_phase(pc, plan)[1])
        self.nbuf = lookahead + 1
        self.inputs = [example.clone() for _ in range(self.nbuf)]
        self.coord_streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(lookahead)]
        cur = torch.cuda.current_stream(dev)
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm), torch.no_grad():
            for _ in range(warmup):   # lazily built state (folded weights, scratch) must exist before capture
                self.fn(self.inputs[0], backbone.coordinate_phase(self.inputs[0]))
        cur.wait_stream(warm)
        torch.cuda.synchronize(dev)
        self.coord_graphs, self.feat_graphs, self.plans, self.outputs = [], [], [], []
        self.cold_start = cold_start if fps_mode != 0 else 0
        self.cold_graphs = []
        for mode in ([fps_mode, 0] if self.cold_start > 0 else [fps_mode]):
            prev_mode = native.set_fps_mode(mode)
            try:
                for k in range(self.nbuf):   # the library's cached scratch of every arena must exist before capture
                    prev_arena = native.set_workspace_arena(1 + k)
                    try:
                        with torch.no_grad():
                            backbone.coordinate_phase(self.inputs[k])
                    finally:
                        native.set_workspace_arena(prev_arena)
            finally:
                native.set_fps_mode(prev_mode)
        torch.cuda.synchronize(dev)
        prev_mode = native.set_fps_mode(fps_mode)
        try:
            for k in range(self.nbuf):
                # every buffer set owns its scratch arena: coordinate phases of different batches overlap in time
                prev_arena = native.set_workspace_arena(1 + k)
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.no_grad(), torch.cuda.graph(g):
                        plan = backbone.coordinate_phase(self.inputs[k])
                finally:
                    native.set_workspace_arena(prev_arena)
                self.coord_graphs.append(g)
                self.plans.append(plan)
        finally:
            native.set_fps_mode(prev_mode)
        if self.cold_start > 0:
            prev_mode = native.set_fps_mode(0)
            try:
                for k in range(self.nbuf):
                    prev_arena = native.set_workspace_arena(1 + k)
                    try:
                        g = torch.cuda.CUDAGraph()
                        with torch.no_grad(), torch.cuda.graph(g):
                            cold = backbone.coordinate_phase(self.inputs[k])
                            for dst, src in zip(_plan_tensors(self.plans[k]), _plan_tensors(cold)):
                                assert dst.shape == src.shape and dst.dtype == src.dtype
                                if dst.data_ptr() != src.data_ptr():   # views of the input buffer are shared as they are
                                    dst.copy_(src)
                    finally:
                        native.set_workspace_arena(prev_arena)
                    self.cold_graphs.append(g)
            finally:
                native.set_fps_mode(prev_mode)
        for k in range(self.nbuf):
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g):
                out = self.fn(self.inputs[k], self.plans[k])
            self.feat_graphs.append(g)
            self.outputs.append(out)
        self.head = 0                        # batches submitted
        self.tail = 0                        # batches completed
        self._ready = [None] * self.nbuf     # coordinate-stream events: plan[k] is complete
        self._done = [None] * self.nbuf      # main-stream events: the feature phase that read buffer set k has finished

    def submit(self, batch: torch.Tensor, after: torch.cuda.Event = None):
        """Stage `batch` (H2D / D2D copy on a coordinate stream) and start its coordinate phase; returns immediately.
        `after`: event that marks a device-resident batch as produced (the coordinate streams do not otherwise wait
        for the caller's stream, that is the point)."""
        assert self.head - self.tail <= self.lookahead, "complete() a batch before submitting more"
        k = self.head % self.nbuf
        st = self.coord_streams[self.head % self.lookahead]
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(st):
            if self._done[k] is not None:
                st.wait_event(self._done[k])
            else:
                st.wait_stream(main)
            if after is not None:
                st.wait_event(after)
            self.inputs[k].copy_(batch, non_blocking=True)
            if self.head - self.tail < self.cold_start:
                self.cold_graphs[k].replay()
            else:
                self.coord_graphs[k].replay()
            ev = torch.cuda.Event()
            ev.record(st)
        self._ready[k] = ev
        self.head += 1

    def complete(self, consume: Callable = None):
        """Runs the feature phase of the oldest submitted batch (graph replay) and returns its output -- or, with
        `consume`, the value of consume(output), called on the stream the feature phase runs on (so a reduction or a
        device-to-host copy of the result is ordered after it without blocking anything else)."""
        assert self.tail < self.head, "nothing submitted"
        k = self.tail % self.nbuf
        cur = torch.cuda.current_stream(self.device)
        st = cur if self.feature_streams is None else self.feature_streams[self.tail % len(self.feature_streams)]
        with torch.cuda.stream(st):
            st.wait_event(self._ready[k])
            self.feat_graphs[k].replay()
            ev = torch.cuda.Event()
            ev.record(st)
            self._done[k] = ev
            self.tail += 1
            out = self.outputs[k]
            return out if consume is None else consume(out)

    def fork(self):
        """Orders the internal feature streams after the work already queued on the caller's stream."""
        if self.feature_streams is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            for st in self.feature_streams:
                st.wait_event(ev)

    def join(self):
        """Orders the caller's stream after everything queued on the internal feature streams."""
        if self.feature_streams is not None:
            cur = torch.cuda.current_stream(self.device)
            for st in self.feature_streams:
                cur.wait_stream(st)
