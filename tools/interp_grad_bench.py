"""three_interpolate_grad at the four feature-propagation shapes of the Stage-1 training step (32 scenes per GPU):
the gather over the inverse stencil (csrc/interp_grad.cu) or, with WS3D_INTERP_GRAD_ATOMIC=1, the atomic scatter.

    python tools/interp_grad_bench.py                        -> one JSON line (ms per call, algorithmic GB/s)
    WS3D_INTERP_GRAD_ATOMIC=1 python tools/interp_grad_bench.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import native, pointnet2_utils, synth  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pts = torch.from_numpy(synth.make_batch(B, 16384)).to(dev)
xyz = [pts[..., :3].contiguous()]
for m in (4096, 1024, 256, 64):
    xyz.append(pointnet2_utils.sample_and_gather(xyz[-1], m)[1])
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
out = {"mode": "atomic scatter" if os.environ.get("WS3D_INTERP_GRAD_ATOMIC") == "1" else "gather over the inverse stencil", "scenes": B,
       "levels": []}
total = 0.0
for lvl, c in ((0, 256), (1, 512), (2, 512), (3, 1024)):        # FP0 .. FP3: channels of the interpolated (deeper) level
    unknown, known = xyz[lvl], xyz[lvl + 1]
    n, m = unknown.shape[1], known.shape[1]
    idx, w = pointnet2_utils.three_nn_weights(unknown, known)
    g = torch.randn(B, c, n, device=dev)
    gp = torch.zeros(B, c, m, device=dev)
    ts = []
    for it in range(8):
        gp.zero_()
        flush.fill_(0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        native.three_interpolate_grad_wrapper(B, c, n, m, g, idx, w, gp)
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    byts = B * (4 * c * n + 24 * n + 8 * c * m)
    total += ms
    out["levels"].append({"fp": lvl, "c": c, "n": n, "m": m, "ms": round(ms, 4), "alg_bytes": byts, "GBps": round(byts / ms / 1e6, 1)})
out["total_ms"] = round(total, 4)
print(json.dumps(out))
