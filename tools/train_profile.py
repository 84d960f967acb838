"""Where a Stage-1 RPN training step spends its GPU time (torch.profiler / kineto, eager launches): top kernels."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import models, synth  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
net = models.RPN().to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=2e-3)
B = 16
pts = torch.from_numpy(synth.make_batch(B, 16384)).to(dev)
g = torch.Generator(device="cpu").manual_seed(0)
cls_label = (torch.rand(B, 16384, generator=g) < 0.05).float().to(dev)
reg_label = torch.randn(B, 16384, 40, generator=g).to(dev)


def step():
    out = net({"pts_input": pts})
    logit = out["rpn_cls"].squeeze(-1)
    p = torch.sigmoid(logit)
    focal = (0.25 * cls_label * (1 - p) ** 2 + 0.75 * (1 - cls_label) * p ** 2) * \
        F.binary_cross_entropy_with_logits(logit, cls_label, reduction="none")
    fg = cls_label.unsqueeze(-1)
    loss = focal.sum() / cls_label.sum().clamp_min(1.0) + \
        (F.smooth_l1_loss(out["rpn_reg"], reg_label, reduction="none") * fg).sum() / fg.sum().clamp_min(1.0)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


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
print(f"GPU time per step: {total / 2e3:.2f} ms (sum of kernels, 2 steps profiled)")
for r in rows[:30]:
    print(f"{r.self_device_time_total / 2e3:8.3f} ms/step  {100 * r.self_device_time_total / total:5.1f} %  x{r.count // 2:<4d} {r.key[:110]}")
