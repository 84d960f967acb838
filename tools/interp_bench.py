"""three_interpolate timings at the FP-layer shapes (B=16)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import native
dev = "cuda:0"; B = 16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (c, m, n) in [(256, 4096, 16384), (512, 1024, 4096), (512, 256, 1024), (1024, 64, 256)]:
    pts = torch.randn(B, c, m, device=dev)
    idx = torch.randint(0, m, (B, n, 3), device=dev, dtype=torch.int32)
    w = torch.rand(B, n, 3, device=dev)
    out = torch.empty(B, c, n, device=dev)
    run = lambda: native.three_interpolate_wrapper(B, c, m, n, pts, idx, w, out)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    ms = float(np.median(ts)); byt = B * (4 * c * m + 24 * n + 4 * c * n)
    ref = (pts.gather(2, idx[..., 0].long().unsqueeze(1).expand(-1, c, -1)) * w[..., 0].unsqueeze(1)
           + pts.gather(2, idx[..., 1].long().unsqueeze(1).expand(-1, c, -1)) * w[..., 1].unsqueeze(1)
           + pts.gather(2, idx[..., 2].long().unsqueeze(1).expand(-1, c, -1)) * w[..., 2].unsqueeze(1))
    print(f"c={c} m={m} n={n}: {ms:.4f} ms  {byt / ms / 1e6:.0f} GB/s  max|err| {float((out - ref).abs().max()):.2e}")
