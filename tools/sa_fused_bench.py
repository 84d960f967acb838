"""Set-abstraction scales at the Stage-1 shapes: the fused kernel (csrc/sa_fused.cu) against the per-layer path
(group_concat + 3 x mlp_layer).  `--one` runs a single fused launch per scale (for ncu).  Writes gpurun_out/sa_fused_bench.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ws3d_b200 import models, pointnet2_utils, synth  # noqa: E402

dev = "cuda:0"


def ev_time(fn, iters=10, warm=3):
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
        l_xyz, l_feat = [xyz], [feat]
        for level, sa in enumerate(model.SA_modules[:3]):
            _, nx = pointnet2_utils.sample_and_gather(l_xyz[-1], sa.npoint)
            x, f = l_xyz[-1], l_feat[-1]

            def run():
                return sa(x, f, new_xyz=nx)[1]

            if one:
                os.environ["WS3D_SA_FUSED"] = "1"
                nf = run()
                torch.cuda.synchronize()
            else:
                os.environ["WS3D_SA_FUSED"] = "1"
                t_f = ev_time(run)
                nf = run()
                os.environ["WS3D_SA_FUSED"] = "0"
                t_u = ev_time(run)
                ref = run()
                err = float((nf - ref).abs().max()) / (float(ref.abs().max()) + 1e-9)
                out[f"SA{level + 1}"] = {"fused_ms": t_f, "per_layer_ms": t_u, "rel_diff": err}
                print(level + 1, out[f"SA{level + 1}"], flush=True)
            l_xyz.append(nx)
            l_feat.append(nf)
    if not one:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sa_fused_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
