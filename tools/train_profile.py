"""Where a Stage-1 RPN training step spends its GPU time (torch.profiler / kineto, eager launches): top kernels.
The step is workloads.RpnTrainStep -- forward in training mode, Gaussian labels on the GPU, get_rpn_loss, backward, Adam --
on this library's training kernels (WS3D_TRAIN_MLP=0 profiles the PyTorch / cuDNN MLP path instead).

    python tools/train_profile.py [scenes_per_gpu]          -> gpurun_out/train_profile.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import workloads  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
step = workloads.RpnTrainStep(B, dev, graph=False, prefetch=False)
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.self_device_time_total)
total = sum(r.self_device_time_total for r in rows)
lines = [f"GPU time per step: {total / 2e3:.2f} ms (sum of kernels, 2 steps profiled, {B} scenes, WS3D_TRAIN_MLP={os.environ.get('WS3D_TRAIN_MLP', '1')})"]
for r in rows[:40]:
    lines.append(f"{r.self_device_time_total / 2e3:8.3f} ms/step  {100 * r.self_device_time_total / total:5.1f} %  x{r.count // 2:<4d} {r.key[:120]}")
print("\n".join(lines))
os.makedirs("gpurun_out", exist_ok=True)
with open(os.path.join("gpurun_out", f"train_profile_mlp{os.environ.get('WS3D_TRAIN_MLP', '1')}.txt"), "w") as f:
    f.write("\n".join(lines) + "\n")
