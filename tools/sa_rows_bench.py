"""The fused SA kernel per scale at the Stage-1 shapes (B = 16): channel-major gather against point-major rows, one launch
at a time on one stream (CUDA events, L2 flushed).  `--one`: a single launch of each variant (the command ncu wraps).
Writes gpurun_out/sa_rows_bench.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ws3d_b200 import fused_mlp, models, pointnet2_utils, synth  # noqa: E402

dev = "cuda:0"


def ev_time(fn, iters=20, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def main():
    one = "--one" in sys.argv
    torch.manual_seed(0)
    model = models.Pointnet2MSG(input_channels=1).to(dev).eval()
    pts = torch.from_numpy(synth.make_batch(16)).to(dev)
    xyz, feat = model._break_up_pc(pts)
    out = {}
    with torch.no_grad():
        rows, x, f = pts, xyz, feat
        for level, sa in enumerate(model.SA_modules[:2]):
            _, nx = pointnet2_utils.sample_and_gather(x, sa.npoint)
            idx = sa._neighbour_indices(x, nx)
            widths = [m[-1].conv.out_channels for m in sa.mlps]
            res = torch.empty((16, sum(widths), sa.npoint), device=dev)
            ld_pm = (3 + sum(widths) + 7) // 8 * 8
            out_pm = torch.empty((16, sa.npoint, ld_pm), device=dev)
            off = 0
            for k, (mlp, ix) in enumerate(zip(sa.mlps, idx)):
                scale = fused_mlp.FusedSAScale(mlp)
                variants = {"channel_major": lambda: scale(x, nx, f, ix, res, off),
                            "rows": lambda: scale(x, nx, f, ix, res, off, rows=rows),
                            "rows+emit": lambda: scale(x, nx, f, ix, res, off, rows=rows, out_pm=out_pm, pm_xyz=(k == 0))}
                for name, fn in variants.items():
                    if one:
                        fn()
                    else:
                        out[f"SA{level + 1}.scale{k}.{name}_ms"] = ev_time(fn)
                        print(f"SA{level + 1} scale {k} {name}: {out[f'SA{level + 1}.scale{k}.{name}_ms']:.4f} ms", flush=True)
                off += widths[k]
            torch.cuda.synchronize()
            rows, x, f = out_pm, nx, res
    if not one:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sa_rows_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
