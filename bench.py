#!/usr/bin/env python
"""bench.py -- the headline measurement of ws3d_b200 (contract in the task statement).

Workload (BASELINE.json configs[1]): full PointNet++-MSG backbone forward (4 SA + 4 FP layers,
tools/cfgs/weaklyRPN.yaml shapes) on a batch of 16 synthetic KITTI-shaped clouds (16384 points x 4
channels) per GPU.  Metric: set-abstraction path throughput in Mpoints/s = clouds x 16384 / time.

  python bench.py [--gpus N --steps K --warmup W]          this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference ...                     the CPU path (oracle port, all host cores)

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM.  `e2e`: the same forward called
with HOST (pinned) input, H2D copy and a D2H read of the per-cloud result checksum inside the timed
region.  `roofline`: the dominant kernel of the step, timed live with CUDA events on the launching
stream.  `cpu_baseline`: the oracle port on the host cores for a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SA-layer Mpoints/sec (PointNet++-MSG backbone forward: 4 SA + 4 FP)"
UNIT = "Mpoints/s"
NPTS = 16384
BATCH = 16
WORKLOAD = ("PointNet++-MSG backbone forward (4 SA + 4 FP, weaklyRPN.yaml shapes), batch 16 synthetic KITTI clouds 16384x4 per GPU "
            "(BASELINE configs[1])")


def measured_traffic(kernel, dims):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/r1_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)
    key = kernel + ":" + ",".join(f"{k}={dims[k]}" for k in sorted(dims))
    return table.get(key)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (rank 0 only)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._th = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._th.join(timeout=6)
        return False

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port).  This is the only place besides tests/ and smoke() that touches oracle/.
def cpu_backbone_throughput(clouds, repeats=1, threads=None):
    import torch

    import oracle
    from oracle import cpu_backbone
    from ws3d_b200 import models, synth
    if threads:
        oracle.set_num_threads(threads)
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = models.Pointnet2MSG(input_channels=1).eval()
    pts = synth.make_batch(clouds)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu_backbone.backbone_forward(model, pts)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return clouds * NPTS / best / 1e6, best, oracle.num_threads()


def run_reference_cpu(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0))
    clouds = BATCH
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_backbone_throughput(1, threads=cores)
    times = []
    for _ in range(args.steps):
        _, dt, thr = cpu_backbone_throughput(clouds, threads=cores)
        times.append(dt)
    ms = float(np.mean(times)) * 1e3
    value = clouds * NPTS / (ms / 1e3) / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clouds_per_step": clouds,
                   "note": "CPU port of the reference algorithms (the reference has no CPU implementation of these ops); each "
                           "step is one pass over the GPU arm's 16-cloud batch (about 2.5 s on 16 cores)"},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{clouds} clouds x {NPTS} points per step, oracle ops (OpenMP) + PyTorch CPU MLPs"},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
ALG_BYTES = {
    # SURVEY.md section 8d, per launch; b = clouds in the launch
    "fps": lambda b, n, m, **k: b * (12 * n + 4 * m),
    "ball_query2": lambda b, n, m, k0, k1, **k: b * (12 * n + 12 * m + 4 * m * (k0 + k1)),
    "group_concat": lambda b, n, m, c, k, **kw: b * (4 * m * k + 12 * n + 4 * c * n + 12 * m + 4 * (3 + c) * m * k),
    "three_nn": lambda b, n, m, **k: b * (12 * n + 12 * m + 24 * n),
    "three_interpolate": lambda b, c, m, n, **k: b * (4 * c * m + 24 * n + 4 * c * n),
    # one fused set-abstraction scale: indices + coordinates + features in, pooled features out (weights are L2-resident)
    "sa_mlp_fused": lambda b, n, m, k, c, c3, **kw: b * (4 * m * k + 12 * n + 12 * m + 4 * c * n + 4 * c3 * m),
    # one shared-MLP layer: read (c1+c2) x cols, write c_out x cols (or cols/pool), read the folded weights once
    "mlp_layer": lambda b, c_out, c_in, cols, pool, **k: 4 * (b * c_in * cols + b * c_out * (cols // pool if pool else cols)
                                                              + c_out * c_in),
}


class OpProfiler:
    """CUDA-event timing of this repo's kernels inside the timed region (same stream)."""

    def __init__(self, torch):
        self.torch = torch
        self.records = []  # (op, dims, start, end)

    def wrap(self, native):
        t = self.torch
        prof = self

        def timed(name, fn, dims_fn):
            def inner(*a):
                s, e = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                s.record()
                r = fn(*a)
                e.record()
                prof.records.append((name, dims_fn(*a), s, e))
                return r
            return inner

        self._orig = {}
        table = {
            "furthest_point_sampling_gather": ("fps", lambda b, n, m, *r: dict(b=b, n=n, m=m)),
            "ball_query2": ("ball_query2", lambda b, n, m, r0, k0, r1, k1, *r: dict(b=b, n=n, m=m, k0=k0, k1=k1)),
            "group_concat": ("group_concat", lambda b, n, m, c, k, *r: dict(b=b, n=n, m=m, c=c, k=k)),
            "three_nn_wrapper": ("three_nn", lambda b, n, m, *r: dict(b=b, n=n, m=m)),
            "three_interpolate_wrapper": ("three_interpolate", lambda b, c, m, n, *r: dict(b=b, c=c, m=m, n=n)),
            "sa_mlp_fused": ("sa_mlp_fused", lambda b, n, m, nsample, c_feat, xyz, new_xyz, features, idx, widths, *r:
                             dict(b=b, n=n, m=m, k=nsample, c=c_feat, c3=widths[2])),
            "mlp_layer": ("mlp_layer", lambda b, c_out, c_out_pad, c1, c2, cols, w, shift, x1, x2, out, relu, pool:
                          dict(b=b, c_out=c_out, c_in=c1 + c2, cols=cols, pool=pool)),
        }
        for attr, (name, dims) in table.items():
            self._orig[attr] = getattr(native, attr)
            setattr(native, attr, timed(name, self._orig[attr], dims))
        self._native = native

    def unwrap(self):
        for attr, fn in self._orig.items():
            setattr(self._native, attr, fn)

    def summarize(self, steps):
        agg = {}
        for name, dims, s, e in self.records:
            ms = s.elapsed_time(e)
            key = (name, tuple(sorted(dims.items())))
            a = agg.setdefault(key, {"ms": 0.0, "launches": 0, "bytes": ALG_BYTES[name](**dims)})
            a["ms"] += ms
            a["launches"] += 1
        out = []
        for (name, dims), a in agg.items():
            avg = a["ms"] / a["launches"]
            out.append({"kernel": name, "dims": dict(dims), "avg_ms": avg, "ms_per_step": a["ms"] / steps,
                        "alg_bytes": a["bytes"], "GBps": a["bytes"] / avg / 1e6})
        out.sort(key=lambda r: -r["ms_per_step"])
        return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from ws3d_b200 import _C, models, native, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = models.Pointnet2MSG(input_channels=1).to(dev).eval()
    # every rank works on its own clouds (scenes shard across GPUs; no data-path collective)
    host = torch.from_numpy(synth.make_batch(BATCH, NPTS, first_scene=rank * BATCH)).pin_memory()
    resident = host.to(dev)

    def step_eager():
        with torch.no_grad():
            return model(resident)[1]

    use_graph = os.environ.get("WS3D_CUDA_GRAPH", "1") != "0"
    if use_graph:
        # the forward (both streams) captured once, replayed per step: same kernels, no per-launch host work
        from ws3d_b200.graphs import CudaGraphRunner
        runner = CudaGraphRunner(lambda x: model(x)[1], resident)

        def step_resident():
            return runner(runner.static_in)

        def step_e2e():
            runner.static_in.copy_(host, non_blocking=True)   # H2D of this step's clouds (pinned source)
            feats = runner(runner.static_in)
            return feats.sum(dim=(1, 2)).cpu()                 # D2H read of the per-cloud checksum (synchronises)
    else:
        step_resident = step_eager

        def step_e2e():
            with torch.no_grad():
                x = host.to(dev, non_blocking=True)
                feats = model(x)[1]
                return feats.sum(dim=(1, 2)).cpu()  # D2H read of the per-cloud checksum (synchronises)

    # Software pipeline (ws3d_b200.graphs.PipelinedBackboneRunner): one replay = level-1 FPS of batch i+1 beside the
    # rest of the forward pass of batch i.  A step still completes exactly one batch of 16 clouds.
    pipelined = use_graph and args.inflight >= 2
    if pipelined:
        from ws3d_b200.graphs import PipelinedBackboneRunner
        native.set_sm_budget(args.sm_budget)
        host_b = torch.from_numpy(synth.make_batch(BATCH, NPTS, first_scene=(world + rank) * BATCH)).pin_memory()
        hosts = [host, host_b]
        pr = PipelinedBackboneRunner(model, resident)
        pr.stage[0].copy_(host)
        pr.stage[1].copy_(host_b)
        pr.prefetch(pr.stage[0])

        def step_pipe():
            return pr.step()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed_region(step_fn, steps, profile):
        prof = None
        if profile:
            prof = OpProfiler(torch)
            prof.wrap(native)
        launches0 = _C.launch_count()
        total = 0.0
        sync_all()
        t_wall = time.perf_counter()
        for _ in range(steps):
            flush.fill_(0)  # L2 flush between timed iterations (outside the event pair)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            step_fn()
            e.record()
            e.synchronize()
            total += s.elapsed_time(e)
        sync_all()
        wall = time.perf_counter() - t_wall
        launches = _C.launch_count() - launches0
        if prof:
            prof.unwrap()
        t = torch.tensor([total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, prof, wall

    def timed_stream_e2e(steps):
        """K pipelined steps from pinned HOST batches, one event pair round all of them: per step an H2D copy of the
        batch after next (copy stream, overlapping the running step), one replay, a checksum kernel and an async
        D2H read of it; the host only ever waits for the PREVIOUS step's result.  L2 is flushed in-stream."""
        pinned = [torch.empty(BATCH, dtype=torch.float32).pin_memory() for _ in range(2)]
        read_ev = [None, None]
        results = []
        pr.prefetch(hosts[0])
        pr.stage_next(hosts[1])
        sync_all()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for j in range(steps):
            flush_small.fill_(0)
            out = pr.step()                       # completes batch j (samples batch j+1 meanwhile)
            pr.stage_next(hosts[j % 2])           # H2D of batch j+2, waits only for the replay that read that buffer
            if read_ev[j % 2] is not None:        # result of step j-2 must have been consumed before its buffer is reused
                read_ev[j % 2].synchronize()
                results.append(float(pinned[j % 2][0]))
            pinned[j % 2].copy_(out.sum(dim=(1, 2)), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            read_ev[j % 2] = ev
        for k in ((steps) % 2, (steps + 1) % 2):
            if read_ev[k] is not None:
                read_ev[k].synchronize()
                results.append(float(pinned[k][0]))
        e.record()
        e.synchronize()
        sync_all()
        t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert len(results) == steps
        return float(t.item())

    flush_small = torch.empty(160 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
        step_eager()
        if pipelined:
            step_pipe()
    if pipelined:
        timed_stream_e2e(max(args.warmup, 3))
    sync_all()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.__enter__()
    ms_total, launches, _, _ = timed_region(step_resident, args.steps, profile=False)
    if use_graph:   # replays do not pass through the C ABI: count the kernels of one eager step instead
        l0 = _C.launch_count()
        step_eager()
        torch.cuda.synchronize()
        launches = (_C.launch_count() - l0) * args.steps
    ms_e2e, _, _, _ = timed_region(step_e2e, args.steps, profile=False)
    ms_single, ms_single_e2e = ms_total, ms_e2e
    if pipelined:
        pr.prefetch(pr.stage[0])
        ms_total, _, _, _ = timed_region(step_pipe, args.steps, profile=False)
        ms_e2e = timed_stream_e2e(args.steps)
    ms_two, ms_two_e2e = ms_total, ms_e2e
    deep = pipelined and args.inflight >= 3
    if deep:
        # The coordinate phase (FPS levels in throughput mode, ball queries, stencils) runs `inflight - 1` batches ahead
        # on its own streams beside the feature phase (ws3d_b200.graphs.StreamedBackboneRunner).  K steps, one event pair.
        from ws3d_b200.graphs import StreamedBackboneRunner
        look = args.inflight - 1
        sr = StreamedBackboneRunner(model, resident, lookahead=look, feature_streams=args.feature_streams)

        def flush_stream(runner, j):   # the stream the feature phase of step j will run on
            fs = runner.feature_streams
            return torch.cuda.current_stream(dev) if fs is None else fs[runner.tail % len(fs)]
        residents = [resident, host_b.to(dev)]

        def timed_streamed(sr, steps, from_host):
            src = hosts if from_host else residents
            pinned = [torch.empty(BATCH, dtype=torch.float32).pin_memory() for _ in range(2)]
            read_ev = [None, None]
            for j in range(look):
                sr.submit(src[j % 2])
            sync_all()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            sr.fork()

            def consume(j):
                def fn(out):
                    if read_ev[j % 2] is not None:            # the pinned slot of step j-2 has been read by now
                        read_ev[j % 2].synchronize()
                    pinned[j % 2].copy_(out.sum(dim=(1, 2)), non_blocking=True)   # per-cloud checksum -> async D2H
                    ev = torch.cuda.Event()
                    ev.record()
                    read_ev[j % 2] = ev
                return fn

            for j in range(steps):
                # (the runner alternates consecutive feature phases between its `feature_streams` internal streams)
                with torch.cuda.stream(flush_stream(sr, j)):
                    flush_small.fill_(0)                      # in-stream L2 flush, inside the timed region
                sr.complete(consume(j) if from_host else None)    # feature phase of batch j (+ result read-back)
                sr.submit(src[(j + look) % 2])                # staging copy + coordinate phase of batch j + look (own stream)
            sr.join()
            e.record()
            e.synchronize()
            for _ in range(look):                         # drain the batches sampled ahead
                sr.complete()
            sync_all()
            t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        timed_streamed(sr, max(args.warmup, 3), False)
        timed_streamed(sr, max(args.warmup, 3), True)
        ms_total = timed_streamed(sr, args.steps, False)
        ms_e2e = timed_streamed(sr, args.steps, True)
    # the same K steps once more with a CUDA-event pair round every launch of this library (per-kernel durations
    # for the roofline entries; kept out of `value` because ~600 extra event records per step cost host time)
    def step_eager_two_phase():   # what the streamed pipeline runs, serially: coordinate phase (throughput FPS), feature phase
        with torch.no_grad():
            prev = native.set_fps_mode(1)
            try:
                plan = model.coordinate_phase(resident)
            finally:
                native.set_fps_mode(prev)
            return model.feature_phase(resident, plan)[1]

    if deep:
        for _ in range(2):
            step_eager_two_phase()
    ms_prof, _, prof, _ = timed_region(step_eager_two_phase if deep else step_eager, args.steps, profile=True)
    # Stage-1 RPN = backbone + the two per-point heads (lib/net/rpn.py:67-81): scenes/s for the metric's second half
    rpn = models.RPN().to(dev).eval()
    rpn.backbone_net = model

    def rpn_heads(pc, first=None):
        o = rpn(pc, first_samples=first)
        return o["rpn_cls"], o["rpn_reg"]

    if pipelined:
        rpn_pr = PipelinedBackboneRunner(model, resident, fn=rpn_heads)
        rpn_pr.stage[0].copy_(host)
        rpn_pr.stage[1].copy_(host_b)
        rpn_pr.prefetch(rpn_pr.stage[0])

        def step_rpn():
            return rpn_pr.step()
    elif use_graph:
        rpn_runner = CudaGraphRunner(rpn_heads, resident)

        def step_rpn():
            return rpn_runner(rpn_runner.static_in)
    else:
        def step_rpn():
            with torch.no_grad():
                return rpn(resident)["rpn_cls"]

    for _ in range(3):
        step_rpn()
    ms_rpn, _, _, _ = timed_region(step_rpn, args.steps, profile=False)
    if deep:
        def rpn_heads_plan(pc, plan):
            o = rpn(pc, plan=plan)
            return o["rpn_cls"], o["rpn_reg"]

        rpn_sr = StreamedBackboneRunner(model, resident, fn=rpn_heads_plan, lookahead=look, feature_streams=args.feature_streams)
        timed_streamed(rpn_sr, 3, False)
        ms_rpn = timed_streamed(rpn_sr, args.steps, False)
    if sampler:
        sampler.__exit__(None, None, None)

    ms_step = ms_total / args.steps
    value = world * BATCH * NPTS / (ms_step / 1e3) / 1e6
    e2e_value = world * BATCH * NPTS / (ms_e2e / args.steps / 1e3) / 1e6
    peak, peak_src = measured_peaks()
    roof = None
    kernels = []
    if prof:
        kernels = prof.summarize(args.steps)
        top = kernels[0]
        roof = {"bound": "hbm", "kernel": top["kernel"], "dims": top["dims"], "achieved": round(top["GBps"], 3), "peak": peak,
                "unit": "GB/s", "frac": round(top["GBps"] / peak, 6), "traffic": measured_traffic(top["kernel"], top["dims"]),
                "peak_source": peak_src, "alg_bytes": int(top["alg_bytes"]),
                "avg_launch_ms": round(top["avg_ms"], 4), "share_of_step": round(top["ms_per_step"] / (ms_prof / args.steps), 4),
                "profiled_ms_per_step": round(ms_prof / args.steps, 4),
                "note": ("FPS is a latency chain of m-1 dependent iterations (SURVEY.md 8d): its HBM fraction is reported for the "
                         "record; us/iteration is the meaningful figure.  In the pipelined modes it runs beside the feature "
                         "phase of other batches (throughput mode: one SM per cloud), so its share of the SERIAL profile pass "
                         "below is not its share of the timed step" if top["kernel"] == "fps" else "")}
        if top["kernel"] == "fps":
            roof["us_per_iteration"] = round(top["avg_ms"] * 1e3 / (top["dims"]["m"] - 1), 4)

    if rank == 0:
        cores = len(os.sched_getaffinity(0))
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, dt, thr = cpu_backbone_throughput(BATCH, repeats=2, threads=cores)
            cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"the full batch ({BATCH} clouds x {NPTS} points), best of 2 passes ({dt:.1f} s each): oracle ops "
                             f"(OpenMP, {thr} threads) + PyTorch CPU MLPs"}
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "clouds_per_gpu": BATCH, "points_per_cloud": NPTS, "l2": ("flushed before every step, in-stream and inside the timed region (160 MiB write > 126 MB L2)" if deep else
                              "flushed between timed iterations (256 MiB write, outside the per-step event pair)"),
                       "mlp": ("tcgen05 TF32, FP32 accumulate: SA1 / SA2 one fused kernel per scale (grouping + 3 layers + max-pool in "
                               "tensor memory), other layers one launch per conv1x1+BN+ReLU[+max-pool]"
                               if torch.backends.cudnn.allow_tf32 else "PyTorch/cuDNN fp32 (TF32 off)"),
                       "launch": ("two CUDA graph replays per step (coordinate phase of batch i+N-1, feature phase of batch i)" if deep
                                  else "one CUDA graph replay per step" if use_graph else "eager launches"),
                       "pipeline": ((f"{args.inflight} batches in flight: the coordinate phase (4 FPS levels in throughput mode = one "
                                     f"SM per cloud, ball queries, interpolation stencils) runs {args.inflight - 1} batches ahead on "
                                     f"its own streams beside the graph-replayed feature phases, which alternate between "
                                     f"{args.feature_streams} stream(s); one batch of 16 clouds completes per step; K steps under "
                                     "one event pair, L2 flushed in-stream") if deep else
                                    ("2 batches in flight: each replay runs level-1 FPS of batch i+1 (high-priority stream) beside "
                                     "the rest of the forward pass of batch i; one batch of 16 clouds completes per step")
                                    if pipelined else "none (one batch in flight)"),
                       "sm_budget_persistent_kernels": args.sm_budget if pipelined else 0,
                       "streams": "two CUDA streams (FPS chain + interpolation stencils run ahead of grouping / MLPs)"
                                  if os.environ.get("WS3D_TWO_STREAMS", "1") != "0" else "single stream",
                       "sharding": "scenes per rank, no data-path collective"},
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(host.numel() * 4) * world,
                    "d2h_bytes_per_step": BATCH * 4 * world, "ms_per_step": round(ms_e2e / args.steps, 4),
                    "note": ("pinned host clouds -> H2D (copy stream) -> pipelined forward -> per-cloud feature checksum -> async D2H, "
                             "K steps streamed under one event pair, L2 flushed in-stream every step" if pipelined else
                             "pinned host cloud -> H2D -> backbone forward -> per-cloud feature checksum -> D2H")},
            "two_in_flight": ({"ms_per_step": round(ms_two / args.steps, 4),
                               "Mpoints_per_s": round(world * BATCH * NPTS / (ms_two / args.steps / 1e3) / 1e6, 3),
                               "e2e_Mpoints_per_s": round(world * BATCH * NPTS / (ms_two_e2e / args.steps / 1e3) / 1e6, 3),
                               "note": "PipelinedBackboneRunner: FPS of batch i+1 inside the same replay as the rest of batch i"}
                              if deep else None),
            "single_batch_latency": {"ms": round(ms_single / args.steps, 4),
                                     "Mpoints_per_s": round(world * BATCH * NPTS / (ms_single / args.steps / 1e3) / 1e6, 3),
                                     "e2e_ms": round(ms_single_e2e / args.steps, 4),
                                     "note": "one batch in flight (graph replay of the two-stream forward), same timing method"},
            "gpu_launches": int(launches),
            "rpn": {"scenes_per_s": round(world * BATCH / (ms_rpn / args.steps / 1e3), 1), "ms_per_step": round(ms_rpn / args.steps, 4),
                    "note": "Stage-1 RPN forward (backbone + cls/reg heads, lib/net/rpn.py:67-81), same batch, inputs resident"
                            + (", same software pipeline as `value`" if pipelined else "")},
            "clocks": sampler.summary() if sampler else None,
        }
        if roof:
            line["roofline"] = roof
            line["kernels"] = [{"kernel": k["kernel"], "dims": k["dims"], "avg_ms": round(k["avg_ms"], 4),
                                "ms_per_step": round(k["ms_per_step"], 4), "GBps": round(k["GBps"], 2),
                                "frac": round(k["GBps"] / peak, 5)} for k in kernels]
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("WS3D_INFLIGHT", "6")),
                    help=">= 3: coordinate phase (FPS, ball queries, stencils) N-1 batches ahead of the feature phase "
                         "(StreamedBackboneRunner); 2: level-1 FPS of the next batch beside the current batch; 1: no pipeline")
    ap.add_argument("--feature-streams", type=int, default=int(os.environ.get("WS3D_FEATURE_STREAMS", "2")),
                    help="streams the feature phases of consecutive batches alternate between (streamed pipeline)")
    ap.add_argument("--sm-budget", type=int, default=int(os.environ.get("WS3D_SM_BUDGET", "100")),
                    help="SMs a persistent MLP kernel spreads over in the pipelined modes (0 = all 148; the coordinate phases of the "
                         "batches ahead hold 3-5 x 16 SMs, so full-width grids would queue behind them: 185.4 / 186.8 / 187.6 Mpoints/s "
                         "at 0 / 72 / 100)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_cpu(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
