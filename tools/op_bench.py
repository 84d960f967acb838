"""Per-op timings at the BASELINE config shapes: B200 ops vs the reference kernels (oracle/_ref).

    gpurun -- 'python tools/op_bench.py --out gpurun_out/op_bench.json'

CUDA events on torch's current stream, 3 warm-ups, L2 flushed (256 MiB write) before every timed
launch, median of `--iters`.  Prints one JSON line per (op, shape, impl).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refmods import load_ref  # noqa: E402

from ws3d_b200 import native, synth  # noqa: E402

dev = "cuda:0"
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    _flush.fill_(1)


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/op_bench.json")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()
    ref = load_ref("pointnet2_cuda")
    refiou, refroi = load_ref("iou3d_cuda"), load_ref("roipool3d_cuda")
    rows = []

    def rec(op, shape, impl, ms, best, bytes_alg=None, extra=None):
        r = {"op": op, "shape": shape, "impl": impl, "ms": round(ms, 4), "ms_best": round(best, 4)}
        if bytes_alg:
            r["alg_GBps"] = round(bytes_alg / ms / 1e6, 1)
        if extra:
            r.update(extra)
        rows.append(r)
        print(json.dumps(r), flush=True)

    B = args.batch
    pts = torch.from_numpy(synth.make_batch(B)).to(dev)
    xyz0 = pts[..., :3].contiguous()
    levels = [(16384, 4096, (0.1, 0.5)), (4096, 1024, (0.5, 1.0)), (1024, 256, (1.0, 2.0)), (256, 64, (2.0, 4.0))]
    chans = [1, 96, 256, 512]
    xyz_l = [xyz0]
    for li, (n, m, radii) in enumerate(levels):
        xyz = xyz_l[-1]
        for b in sorted({1, B}):
            x = xyz[:b].contiguous()
            shape = f"B{b} N{n} M{m}"
            idx = torch.empty((b, m), dtype=torch.int32, device=dev)
            temp = torch.empty((b, n), device=dev)

            def run_mine():
                temp.fill_(1e10)
                native.furthest_point_sampling_wrapper(b, n, m, x, temp, idx)
            ms, best = timeit(run_mine, args.iters)
            rec("fps", shape, "b200", ms, best, b * (12 * n + 4 * m), {"us_per_iter": round(ms * 1e3 / max(1, m - 1), 4)})
            if ref is not None:
                ridx = torch.empty_like(idx)

                def run_ref():
                    temp.fill_(1e10)
                    ref.furthest_point_sampling_wrapper(b, n, m, x, temp, ridx)
                ms, best = timeit(run_ref, max(3, args.iters // 3))
                rec("fps", shape, "reference", ms, best, b * (12 * n + 4 * m), {"us_per_iter": round(ms * 1e3 / max(1, m - 1), 4)})
                assert torch.equal(idx, ridx)
            # (decomposition sweeps: tools/fps_sweep.py, one subprocess per setting -- the library reads its
            #  WS3D_FPS_* overrides once per process)
        # full-batch sample for the next level
        idxB = torch.empty((B, m), dtype=torch.int32, device=dev)
        new_xyz = torch.empty((B, m, 3), device=dev)
        native.furthest_point_sampling_gather(B, n, m, xyz, None, idxB, new_xyz)
        c = chans[li]
        feat = torch.randn((B, c, n), device=dev)
        shape = f"B{B} N{n} M{m}"
        for r, k in zip(radii, (16, 32)):
            bi = torch.zeros((B, m, k), dtype=torch.int32, device=dev)
            ms, best = timeit(lambda: native.ball_query_wrapper(B, n, m, r, k, new_xyz, xyz, bi), args.iters)
            rec("ball_query", f"{shape} r{r} K{k}", "b200", ms, best, B * (12 * n + 12 * m + 4 * m * k))
            if ref is not None:
                rbi = torch.zeros_like(bi)
                ms, best = timeit(lambda: ref.ball_query_wrapper(B, n, m, r, k, new_xyz, xyz, rbi), args.iters)
                rec("ball_query", f"{shape} r{r} K{k}", "reference", ms, best, B * (12 * n + 12 * m + 4 * m * k))
                assert torch.equal(bi, rbi)
            out = torch.empty((B, 3 + c, m, k), device=dev)
            ms, best = timeit(lambda: native.group_concat(B, n, m, c, k, True, xyz, new_xyz, feat, bi, out), args.iters)
            rec("group_concat", f"{shape} C{c} K{k}", "b200", ms, best, B * (4 * m * k + 12 * n + 4 * c * n + 12 * m + 4 * (3 + c) * m * k))
            if ref is not None:
                xyz_t = xyz.transpose(1, 2).contiguous()
                g1, g2 = torch.empty((B, 3, m, k), device=dev), torch.empty((B, c, m, k), device=dev)

                def ref_group():
                    xt = xyz.transpose(1, 2).contiguous()
                    ref.group_points_wrapper(B, 3, n, m, k, xt, bi, g1)
                    g1.sub_(new_xyz.transpose(1, 2).unsqueeze(-1))
                    ref.group_points_wrapper(B, c, n, m, k, feat, bi, g2)
                    return torch.cat([g1, g2], dim=1)
                ms, best = timeit(ref_group, args.iters)
                rec("group_concat", f"{shape} C{c} K{k}", "reference(+torch glue)", ms, best,
                    B * (4 * m * k + 12 * n + 4 * c * n + 12 * m + 4 * (3 + c) * m * k))
        b0 = torch.zeros((B, m, 16), dtype=torch.int32, device=dev)
        b1 = torch.zeros((B, m, 32), dtype=torch.int32, device=dev)
        ms, best = timeit(lambda: native.ball_query2(B, n, m, radii[0], 16, radii[1], 32, new_xyz, xyz, b0, b1), args.iters)
        rec("ball_query2", f"{shape} r{radii}", "b200", ms, best, B * (12 * n + 12 * m + 4 * m * 48))
        # feature propagation at this level: unknown = level points, known = sampled points
        d2 = torch.empty((B, n, 3), device=dev)
        ni = torch.empty((B, n, 3), dtype=torch.int32, device=dev)
        ms, best = timeit(lambda: native.three_nn_wrapper(B, n, m, xyz, new_xyz, d2, ni), args.iters)
        rec("three_nn", f"B{B} n{n} m{m}", "b200", ms, best, B * (12 * n + 12 * m + 24 * n))
        if ref is not None:
            rd2, rni = torch.empty_like(d2), torch.empty_like(ni)
            ms, best = timeit(lambda: ref.three_nn_wrapper(B, n, m, xyz, new_xyz, rd2, rni), args.iters)
            rec("three_nn", f"B{B} n{n} m{m}", "reference", ms, best, B * (12 * n + 12 * m + 24 * n))
            assert torch.equal(ni, rni)
        cf = [256, 512, 512, 1024][li]
        kf = torch.randn((B, cf, m), device=dev)
        w = torch.rand((B, n, 3), device=dev)
        io = torch.empty((B, cf, n), device=dev)
        ms, best = timeit(lambda: native.three_interpolate_wrapper(B, cf, m, n, kf, ni, w, io), args.iters)
        rec("three_interpolate", f"B{B} C{cf} m{m} n{n}", "b200", ms, best, B * (4 * cf * m + 24 * n + 4 * cf * n))
        if ref is not None:
            rio = torch.empty_like(io)
            ms, best = timeit(lambda: ref.three_interpolate_wrapper(B, cf, m, n, kf, ni, w, rio), args.iters)
            rec("three_interpolate", f"B{B} C{cf} m{m} n{n}", "reference", ms, best, B * (4 * cf * m + 24 * n + 4 * cf * n))
        xyz_l.append(new_xyz)

    # ---- config 4: iou3d + roipool3d on one scene
    scene = synth.make_scene(0)
    for nb in (2048, 16384):
        bx3 = synth.make_boxes(scene[:, :3], nb)
        bev = torch.from_numpy(synth.boxes3d_to_bev(bx3)).to(dev)
        scores = torch.rand(nb, device=dev)
        order = scores.sort(descending=True)[1]
        sb = bev[order].contiguous()
        ans = torch.zeros((nb, nb), device=dev)
        ms, best = timeit(lambda: native.boxes_iou_bev_gpu(bev, bev, ans), 5)
        rec("boxes_iou_bev", f"N{nb}", "b200", ms, best, 40 * nb + 4 * nb * nb, {"Mpairs_per_s": round(nb * nb / ms / 1e3, 1)})
        if refiou is not None:
            rans = torch.zeros_like(ans)
            ms, best = timeit(lambda: refiou.boxes_iou_bev_gpu(bev, bev, rans), 3)
            rec("boxes_iou_bev", f"N{nb}", "reference", ms, best, 40 * nb + 4 * nb * nb, {"Mpairs_per_s": round(nb * nb / ms / 1e3, 1)})
            rec("boxes_iou_bev_equal", f"N{nb}", "check", 0, 0, None, {"bit_exact": bool(torch.equal(ans, rans))})
        for th in (0.85, 0.1):
            kb = torch.zeros(nb, dtype=torch.int64)
            ms, best = timeit(lambda: native.nms_gpu(sb, kb, th), 5)
            num = native.nms_gpu(sb, kb, th)
            rec("nms_gpu(host keep)", f"N{nb} th{th}", "b200", ms, best, None, {"kept": num})
            ms, best = timeit(lambda: native.nms_device(sb, th), 5)
            rec("nms_device", f"N{nb} th{th}", "b200", ms, best, None)
            if refiou is not None:
                rkb = torch.zeros(nb, dtype=torch.int64)
                ms, best = timeit(lambda: refiou.nms_gpu(sb, rkb, th), 3)
                rnum = refiou.nms_gpu(sb, rkb, th)
                rec("nms_gpu(host keep)", f"N{nb} th{th}", "reference", ms, best, None,
                    {"kept": rnum, "keep_equal": bool(rnum == num and torch.equal(rkb[:rnum], kb[:num]))})
    for nb, c in ((16384, 1), (16384, 128), (512, 128)):
        bx3 = torch.from_numpy(synth.make_boxes(scene[:, :3], nb)).to(dev)[None].contiguous()
        x = torch.from_numpy(scene[None, :, :3].copy()).to(dev)
        f = torch.randn((1, 16384, c), device=dev)
        pooled = torch.zeros((1, nb, 512, 3 + c), device=dev)
        flag = torch.zeros((1, nb), dtype=torch.int32, device=dev)
        byts = 12 * 16384 + 4 * c * 16384 + 32 * nb + 4 * nb * 512 * (3 + c)
        ms, best = timeit(lambda: native.roipool3d_forward(x, bx3, f, pooled, flag), 5)
        rec("roipool3d", f"M{nb} C{c} S512", "b200", ms, best, byts, {"empty": int(flag.sum())})
        if refroi is not None and nb * 16384 * 4 < (8 << 30):
            rp, rf = torch.zeros_like(pooled), torch.zeros_like(flag)
            ms, best = timeit(lambda: refroi.forward(x, bx3, f, rp, rf), 3)
            rec("roipool3d", f"M{nb} C{c} S512", "reference", ms, best, byts,
                {"equal": bool(torch.equal(rp, pooled) and torch.equal(rf, flag))})
        del pooled
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
