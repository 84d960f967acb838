"""Timings for the SURVEY section-8 "next" rows f2 / f3 on the GPU box, next to the reference formulation run in
torch on the same GPU (the reference's full fg x fg IoU3D matrices + diagonal gather; its Python radius-NMS loop;
its dense distance-matrix cylinder crop).  Writes gpurun_out/next_rows_bench.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ws3d_b200 import iou3d_utils, native, proposal_utils, synth  # noqa: E402

dev = "cuda:0"


def ev_time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def dist(a, b):
    return torch.sqrt(torch.sum((a[None, :] - b[:, None]) ** 2, dim=2))


def main():
    out = {}
    pts = synth.make_scene(3)
    rng = np.random.default_rng(0)
    for n in (256, 4096, 16384):
        a = synth.make_boxes(pts[:, :3], n, seed=n)
        b = a + rng.normal(0, 0.15, a.shape).astype(np.float32)
        ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)

        def full():
            i2, i3 = iou3d_utils.boxes_iou3d_gpu(ta, tb)
            return torch.diagonal(i3)

        rec = {"aligned_ms": ev_time(lambda: iou3d_utils.boxes_iou3d_aligned(ta, tb)),
               "full_matrix_diag_ms": ev_time(full, iters=5)}
        rec["pairs_per_s_aligned"] = n / rec["aligned_ms"] * 1e3
        out[f"iou3d_aligned_n{n}"] = rec
        print(n, rec, flush=True)
    for p in (900, 4096, 16384):
        cen = (pts[rng.integers(0, 16384, p)][:, [0, 2]] + rng.normal(0, 0.25, (p, 2))).astype(np.float32)
        sc = rng.uniform(0, 1, p).astype(np.float32)
        tc, ts_ = torch.from_numpy(cen).to(dev), torch.from_numpy(sc).to(dev)
        rec = {"radius_nms_ms": ev_time(lambda: proposal_utils.radius_nms(tc, ts_, 0.3))}
        order = torch.argsort(-ts_)
        sorted_c = tc[order].contiguous()
        rec["radius_nms_kernels_only_ms"] = ev_time(lambda: native.radius_nms_device(sorted_c, 0.3))
        if p <= 900:   # the script's Python loop (one device sync per candidate)
            def loop():
                rois = tc[torch.argsort(-ts_)]
                keep_id = [0]
                d = dist(rois, rois)
                for i in range(1, rois.shape[0]):
                    if torch.min(d[keep_id, i], dim=-1)[0] > 0.3:
                        keep_id.append(i)
                return keep_id
            t0 = time.perf_counter(); kid = loop(); torch.cuda.synchronize()
            rec["reference_python_loop_ms"] = (time.perf_counter() - t0) * 1e3
            rec["kept"] = len(kid)
        out[f"radius_nms_p{p}"] = rec
        print(p, rec, flush=True)
        keep = proposal_utils.radius_nms(tc, ts_, 0.3)
        centres = tc[keep][:512].contiguous()
        tp = torch.from_numpy(pts).to(dev)

        def dense():
            d = dist(centres, tp[:, [0, 2]])
            any_ = torch.min(d, dim=-1)[0] < 4.0
            return any_, (d < 4.0)

        rec2 = {"centres": int(centres.shape[0]),
                "cylinder_crop_ms": ev_time(lambda: proposal_utils.cylinder_crop(tp, centres, 4.0, cap=2048)),
                "reference_dense_matrix_ms": ev_time(dense)}
        out[f"cylinder_crop_p{p}"] = rec2
        print(p, rec2, flush=True)
    # f4: Gaussian RPN labels for a batch of 16 scenes x 16384 points x 24 boxes, next to the reference's per-sample numpy
    from ws3d_b200 import label_utils
    scenes = np.stack([synth.make_scene(40 + k)[:, :3] for k in range(16)])
    gts = np.stack([synth.make_boxes(scenes[k], 24, seed=k) for k in range(16)])
    tp, tg = torch.from_numpy(scenes).to(dev), torch.from_numpy(gts).to(dev)
    rec = {"gpu_batch16_ms": ev_time(lambda: label_utils.generate_gaussian_training_labels(tp, tg))}
    t0 = time.perf_counter()
    for k in range(16):   # numpy restatement of kitti_rcnn_dataset.py:529-573 (same operations, one scene at a time, host)
        d = np.sqrt((scenes[k][:, None, 0] - gts[k][None, :, 0]) ** 2 + ((scenes[k][:, 1] * 0.707) ** 2)[:, None]
                    + (scenes[k][:, None, 2] - gts[k][None, :, 2]) ** 2)
        m = np.minimum(100.0, np.clip(d - 0.7, 0, 100).min(1))
        _ = np.exp(-0.5 * m.astype(np.float64) ** 2 / 1.5)
        _ = d.argmin(1)
    rec["numpy_host_batch16_ms"] = (time.perf_counter() - t0) * 1e3
    out["gaussian_labels_b16_n16384_g24"] = rec
    print(rec, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "next_rows_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
