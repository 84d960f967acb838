"""Whole-backbone comparison on one GPU: this repo's fused B200 path vs the SAME PyTorch modules
driven by the reference's own kernels (oracle/_ref, recompiled for sm_100a) in the reference's
unfused sequencing (pointnet2_modules.py:19-55, pointnet2_utils.py:241-264).

    gpurun -- 'python tools/backbone_compare.py --out gpurun_out/backbone_compare.json'

Test/benchmark infrastructure only (it loads oracle/_ref); bench.py never does this.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_backbone import RefOps, load_ref, ref_forward  # noqa: E402

from ws3d_b200 import models, synth  # noqa: E402

dev = "cuda:0"


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/backbone_compare.json")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    ref = load_ref("pointnet2_cuda")
    assert ref is not None, "oracle/_ref not built"
    torch.manual_seed(0)
    model = models.Pointnet2MSG(input_channels=1).to(dev).eval()
    pc = torch.from_numpy(synth.make_batch(args.batch)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ops = RefOps(ref)
    res = {"batch": args.batch}
    with torch.no_grad():
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            tag = "tf32" if tf32 else "fp32"
            res[f"b200_ms_{tag}"] = timeit(lambda: model(pc), args.iters, flush)
            res[f"reference_kernels_ms_{tag}"] = timeit(lambda: ref_forward(model, ops, pc), max(3, args.iters // 2), flush)
        a = model(pc)[1]
        b = ref_forward(model, ops, pc)[1]
        res["max_abs_diff_fp32"] = float((a - b).abs().max())
        res["equal_fp32"] = bool(torch.equal(a, b))
    res["Mpoints_per_s_b200_tf32"] = args.batch * 16384 / res["b200_ms_tf32"] / 1e3
    res["Mpoints_per_s_reference_kernels_tf32"] = args.batch * 16384 / res["reference_kernels_ms_tf32"] / 1e3
    print(json.dumps(res))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
