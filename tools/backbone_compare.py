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
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refmods import load_ref  # noqa: E402

from ws3d_b200 import models, synth  # noqa: E402

dev = "cuda:0"


class RefOps:
    """The reference wrappers' call pattern (pointnet2_utils.py) on the reference extension module."""

    def __init__(self, mod):
        self.m = mod

    def fps(self, xyz, npoint):
        B, N, _ = xyz.shape
        out = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, device=xyz.device)
        self.m.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, out)
        return out

    def gather(self, feats, idx):
        B, C, N = feats.shape
        out = torch.empty((B, C, idx.shape[1]), device=feats.device)
        self.m.gather_points_wrapper(B, C, N, idx.shape[1], feats, idx, out)
        return out

    def ball_query(self, r, k, xyz, new_xyz):
        B, N, _ = xyz.shape
        idx = torch.zeros((B, new_xyz.shape[1], k), dtype=torch.int32, device=xyz.device)
        self.m.ball_query_wrapper(B, N, new_xyz.shape[1], r, k, new_xyz, xyz, idx)
        return idx

    def group(self, feats, idx):
        B, C, N = feats.shape
        out = torch.empty((B, C, idx.shape[1], idx.shape[2]), device=feats.device)
        self.m.group_points_wrapper(B, C, N, idx.shape[1], idx.shape[2], feats, idx, out)
        return out

    def three_nn(self, unknown, known):
        B, N, _ = unknown.shape
        d2 = torch.empty((B, N, 3), device=unknown.device)
        idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
        self.m.three_nn_wrapper(B, N, known.shape[1], unknown, known, d2, idx)
        return torch.sqrt(d2), idx

    def interpolate(self, feats, idx, w):
        B, c, m = feats.shape
        out = torch.empty((B, c, idx.shape[1]), device=feats.device)
        self.m.three_interpolate_wrapper(B, c, m, idx.shape[1], feats, idx, w, out)
        return out


def ref_forward(model, ops, pc):
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
    l_xyz, l_f = [xyz], [feats]
    for sa in model.SA_modules:
        x, f = l_xyz[-1], l_f[-1]
        xt = x.transpose(1, 2).contiguous()
        new_xyz = ops.gather(xt, ops.fps(x, sa.npoint)).transpose(1, 2).contiguous()
        outs = []
        for g, mlp in zip(sa.groupers, sa.mlps):
            idx = ops.ball_query(g.radius, g.nsample, x, new_xyz)
            gx = ops.group(x.transpose(1, 2).contiguous(), idx)
            gx -= new_xyz.transpose(1, 2).unsqueeze(-1)
            y = mlp(torch.cat([gx, ops.group(f, idx)], dim=1))
            outs.append(F.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1))
        l_xyz.append(new_xyz)
        l_f.append(torch.cat(outs, dim=1))
    for i in range(-1, -(len(model.FP_modules) + 1), -1):
        dist, idx = ops.three_nn(l_xyz[i - 1], l_xyz[i])
        recip = 1.0 / (dist + 1e-8)
        w = recip / torch.sum(recip, dim=2, keepdim=True)
        interp = ops.interpolate(l_f[i], idx, w)
        nf = torch.cat([interp, l_f[i - 1]], dim=1) if l_f[i - 1] is not None else interp
        l_f[i - 1] = model.FP_modules[i].mlp(nf.unsqueeze(-1)).squeeze(-1)
    return l_xyz[0], l_f[0]


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
