"""Stage-2 (RCNN) set-abstraction stack at BASELINE configs[4] shapes: 512 pooled proposals per scene, 512 points
per proposal, 128 feature channels (tools/cfgs/weaklyRCNN.yaml:60-77, lib/net/rcnn_net.py:40-58).

    python tools/stage2_bench.py [--scenes 1] [--steps 10]      (torchrun for several GPUs: proposals shard per rank)

Forward only, eval mode (tensor-core MLP layers), inputs resident.  Prints one JSON line: proposals/s.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import _C  # noqa: E402
from ws3d_b200.pointnet2_modules import PointnetSAModule, sa_stack_forward  # noqa: E402

SA = {"NPOINTS": [256, 128, 32, -1], "RADIUS": [0.2, 0.4, 1.0, 100.0], "NSAMPLE": [16, 32, 64, 64],
      "MLPS": [[128, 128, 128], [128, 128, 128], [128, 128, 256], [256, 256, 512]]}


class Stage2SA(nn.Module):
    def __init__(self, channel_in=128):
        super().__init__()
        self.SA_modules = nn.ModuleList()
        for k in range(len(SA["NPOINTS"])):
            npoint = SA["NPOINTS"][k] if SA["NPOINTS"][k] != -1 else None
            self.SA_modules.append(PointnetSAModule(npoint=npoint, radius=SA["RADIUS"][k], nsample=SA["NSAMPLE"][k],
                                                    mlp=[channel_in] + SA["MLPS"][k], use_xyz=True, bn=True))
            channel_in = SA["MLPS"][k][-1]

    def forward(self, xyz, features):
        return sa_stack_forward(self.SA_modules, xyz, features)[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=1, help="scenes per GPU (512 proposals each)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--graph", type=int, default=0, help="1: replay the forward as one CUDA graph (no per-launch host work)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = Stage2SA().to(dev).eval()
    B = 512 * args.scenes
    rng = np.random.default_rng(1234 + rank)
    # canonical-frame proposal crops: N(0, (1.2, 0.6, 2.2))
    xyz = torch.from_numpy((rng.normal(0, 1, (B, 512, 3)) * np.array([1.2, 0.6, 2.2])).astype(np.float32)).to(dev)
    feats = torch.randn(B, 128, 512, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(3):
            out = model(xyz, feats)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = _C.launch_count()
        total = 0.0
        run = lambda: model(xyz, feats)                                                # noqa: E731
        if args.graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                model(xyz, feats)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = model(xyz, feats)
            run = lambda: (graph.replay(), static_out)[1]                              # noqa: E731
        for _ in range(args.steps):
            flush.fill_(0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = run()
            e.record()
            e.synchronize()
            total += s.elapsed_time(e)
    t = torch.tensor([total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    if rank == 0:
        print(json.dumps({"metric": "Stage-2 SA stack proposals/sec (4 SA levels, 512 pts x 128 ch per proposal)", "value": round(world * B / (ms / 1e3), 1),
                          "unit": "proposals/s", "n_gpus": world, "steps": args.steps, "ms_per_step": round(ms, 3), "scaling": "weak",
                          "config": {"proposals_per_gpu": B, "points_per_proposal": 512, "out_shape": list(out.shape)},
                          "launch": "graph replay" if args.graph else "eager", "gpu_launches": int(_C.launch_count() - l0)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
