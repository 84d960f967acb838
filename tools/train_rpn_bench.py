"""Stage-1 RPN training step (BASELINE configs[2]): forward + loss + backward + Adam, one rank per GPU.

    python tools/train_rpn_bench.py --batch 16 --steps 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_rpn_bench.py ...

Scenes shard across ranks (no data-path collective); the only exchange is DDP's gradient all-reduce over NCCL,
as in the reference's DataParallel training (tools/train_rpn.py) -- BatchNorm statistics stay per replica.
The training path uses the PyTorch MLPs (batch statistics, autograd) on top of this repo's ops and their
gradient kernels; the tcgen05 MLP kernel is inference-only.  The loss is a stand-in with the reference's
tensor shapes (sigmoid focal loss on the per-point score, smooth-L1 on the 40 bin/residual channels of the
foreground points; lib/net/train_functions.py:60-115 builds the real one from the same two tensors).
Prints one JSON line on rank 0: scenes/s = world x batch / step time (CUDA events, max over ranks).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import _C, models, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16, help="scenes per GPU")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--graph", type=int, default=1, help="1: replay the whole training step (forward, loss, backward, Adam) as one "
                                                         "CUDA graph (single GPU); 0: eager launches")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = models.RPN().to(dev).train()
    n_param_bytes = sum(p.numel() for p in net.parameters()) * 4
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    use_graph = bool(args.graph) and world == 1
    opt = torch.optim.Adam(model.parameters(), lr=2e-3, capturable=use_graph)
    pts = torch.from_numpy(synth.make_batch(args.batch, 16384, first_scene=rank * args.batch)).to(dev)
    g = torch.Generator(device="cpu").manual_seed(rank)
    cls_label = (torch.rand(args.batch, 16384, generator=g) < 0.05).float().to(dev)
    reg_label = torch.randn(args.batch, 16384, 40, generator=g).to(dev)

    def step():
        out = model({"pts_input": pts})
        logit = out["rpn_cls"].squeeze(-1)
        p = torch.sigmoid(logit)
        focal = (0.25 * cls_label * (1 - p) ** 2 + 0.75 * (1 - cls_label) * p ** 2) * \
            F.binary_cross_entropy_with_logits(logit, cls_label, reduction="none")
        loss_cls = focal.sum() / cls_label.sum().clamp_min(1.0)
        fg = cls_label.unsqueeze(-1)
        loss_reg = (F.smooth_l1_loss(out["rpn_reg"], reg_label, reduction="none") * fg).sum() / fg.sum().clamp_min(1.0)
        loss = loss_cls + loss_reg
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    if use_graph:
        # ~2000 launches per step (this library's ops + cuDNN / ATen kernels of the MLPs, their backward and Adam): eager,
        # the host cannot issue them as fast as the GPU retires them.  The step has static shapes, so it is captured once.
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(args.warmup, 3)):
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = step()
        eager_step = step

        def step():   # noqa: F811
            graph.replay()
            return static_loss
    for _ in range(max(args.warmup, 3)):
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
    ms = float(t.item()) / args.steps
    if rank == 0:
        print(json.dumps({"metric": "Stage-1 RPN training scenes/sec (forward + loss + backward + Adam)", "value": round(world * args.batch / (ms / 1e3), 1),
                          "unit": "scenes/s", "n_gpus": world, "steps": args.steps, "ms_per_step": round(ms, 3), "scaling": "weak",
                          "config": {"scenes_per_gpu": args.batch, "points_per_scene": 16384, "optimizer": "Adam", "collective":
                                     "DDP gradient all-reduce (NCCL), %.2f MB fp32" % (n_param_bytes / 1e6) if world > 1 else "none (1 GPU)",
                                     "mlp": "PyTorch/cuDNN (training: batch-norm statistics, autograd)",
                                     "launch": "one CUDA graph replay per training step" if use_graph else "eager launches"},
                          "final_loss": round(float(loss), 4), "gpu_launches": int(_C.launch_count() - l0)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
