"""three_interpolate at the backbone's FP shapes (B = 16): plain and with the affine epilogue (FP0 after the first
convolution moved in front of it: 128 channels).  Prints ms and algorithmic GB/s; checks plain == torch gather formula."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import native, pointnet2_utils  # noqa: E402

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def ev(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        flush.fill_(0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


torch.manual_seed(0)
B = 16
for (c, m, n, affine) in [(128, 4096, 16384, True), (256, 4096, 16384, False), (512, 1024, 4096, False), (512, 256, 1024, False),
                          (1024, 64, 256, False)]:
    pts = torch.randn(B, c, m, device=dev)
    idx = torch.randint(0, m, (B, n, 3), device=dev, dtype=torch.int32)
    w = torch.rand(B, n, 3, device=dev)
    w = w / w.sum(-1, keepdim=True)
    out = torch.empty(B, c, n, device=dev)
    if affine:
        sc, row, sh = torch.randn(c, device=dev), torch.randn(B, n, device=dev), torch.randn(c, device=dev)
        ms = ev(lambda: native.three_interpolate_affine(B, c, m, n, pts, idx, w, sc, row, sh, 3, out))
    else:
        ms = ev(lambda: native.three_interpolate_wrapper(B, c, m, n, pts, idx, w, out))
        want = sum(torch.gather(pts, 2, idx[..., k].long().unsqueeze(1).expand(-1, c, -1)) * w[..., k].unsqueeze(1) for k in (1, 0, 2))
        assert float((out - want).abs().max()) < 1e-4
    byt = B * (4 * c * m + 24 * n + 4 * c * n)
    print(f"c={c} m={m} n={n} affine={affine}: {ms:.4f} ms  {byt / ms / 1e6:.0f} GB/s", flush=True)
