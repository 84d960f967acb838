"""Stage-1 RPN training step (BASELINE configs[2]) from the command line: workloads.RpnTrainStep, one rank per GPU.

    python tools/train_rpn_bench.py --batch 32 --steps 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_rpn_bench.py ...

The step = forward in training mode on this library's training layers (ws3d_b200/train_mlp.py), Gaussian labels on the
GPU, get_rpn_loss, backward, ONE averaged all-reduce of all gradients over NCCL (sharding.FlatGradients; BatchNorm
statistics stay per replica as under the reference's DataParallel, tools/train_rpn.py), Adam -- replayed as one CUDA graph.
Prints one JSON line on rank 0: scenes/s = world x batch / step time (CUDA events, max over ranks), the all-reduce share
(same step without the collective) and whether the replicas still hold identical parameters.  bench.py runs the same
thing as its `configs["3"]` record.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import _C, train_mlp, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32, help="scenes per GPU")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--prefetch", type=int, default=1, help="1: the next batch's coordinate phase runs beside the current step (two steps per call)")
    ap.add_argument("--graph", type=int, default=1, help="1: one CUDA graph replay per step (collective included); 0: eager launches")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def timed(step):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        l0 = _C.launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            loss = step()
        e.record()
        e.synchronize()
        t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / (args.steps * step.steps_per_call), float(loss.detach()), int(_C.launch_count() - l0)

    step = workloads.RpnTrainStep(args.batch, dev, world, rank, graph=bool(args.graph), prefetch=bool(args.prefetch))
    ms, loss, launches = timed(step)
    rec = {"metric": "Stage-1 RPN training scenes/sec (forward + labels + loss + backward + gradient all-reduce + Adam)",
           "value": round(world * args.batch / (ms / 1e3), 1), "unit": "scenes/s", "n_gpus": world, "steps": args.steps * step.steps_per_call,
           "ms_per_step": round(ms, 3), "scaling": "weak", "final_loss": round(loss, 4), "gpu_launches_eager": launches,
           "config": {"prefetch_next_coordinate_phase": bool(args.prefetch), "scenes_per_gpu": args.batch, "points_per_scene": 16384, "optimizer": "Adam", "mlp": train_mlp.DESCRIPTION,
                      "launch": "one CUDA graph replay per training step" if step.graphed else "eager launches",
                      "collective": ("one averaged all-reduce (NCCL) of %.2f MB fp32 per step, inside the graph" % (step.param_bytes / 1e6))
                      if world > 1 else "none (1 GPU)"}}
    if world > 1:
        probe = torch.stack([p.detach().double().sum() for p in step.net.parameters()]).sum().reshape(1)
        lo, hi = probe.clone(), probe.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        rec["replicas_in_sync"] = bool(float(hi - lo) <= 1e-9 * max(1.0, abs(float(hi))))
        del step
        alone = workloads.RpnTrainStep(args.batch, dev, world, rank, graph=bool(args.graph), exchange=False, prefetch=bool(args.prefetch))
        ms_ns, _, _ = timed(alone)
        rec["allreduce"] = {"ms_per_step_without": round(ms_ns, 3), "share_of_step": round(max(0.0, 1.0 - ms_ns / ms), 4)}
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
