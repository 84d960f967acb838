"""Generate tests/golden/*.npz by running the REFERENCE's own kernels (oracle/_ref, built from the
unmodified sources by oracle/build_ref.sh) on a GPU.  Run on the B200 box:

    gpurun -- 'python tools/make_golden.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz to tests/golden/.  Each file holds seeded inputs and the
reference outputs; tests/test_oracle_golden.py pins the CPU oracle to them (no GPU needed) and the
GPU tests pin the CUDA path to them.  Sizes are small on purpose (fixtures are committed).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refmods import load_ref  # noqa: E402

from ws3d_b200 import synth  # noqa: E402

dev = "cuda:0"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    p2, iou, roi = load_ref("pointnet2_cuda"), load_ref("iou3d_cuda"), load_ref("roipool3d_cuda")
    assert p2 and iou and roi, "oracle/_ref is not built"
    rng = np.random.default_rng(20261017)
    meta = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}

    # ---- FPS: scene-like, uniform, tie-heavy integer grid, non power-of-two n
    fps = {}
    for name, (b, n, m, kind) in {"scene": (1, 4096, 1024, "scene"), "uniform": (2, 1024, 256, "u"), "grid": (2, 600, 200, "g"),
                                  "small": (1, 100, 100, "g"), "big": (1, 16384, 512, "scene")}.items():
        if kind == "scene":
            xyz = synth.make_batch(b, n)[..., :3].copy()
        elif kind == "u":
            xyz = rng.uniform(-5, 5, (b, n, 3)).astype(np.float32)
        else:
            xyz = rng.integers(-3, 4, (b, n, 3)).astype(np.float32)
        temp = torch.full((b, n), 1e10, device=dev)
        idx = torch.empty((b, m), dtype=torch.int32, device=dev)
        p2.furthest_point_sampling_wrapper(b, n, m, t(xyz), temp, idx)
        fps[name + "_xyz"], fps[name + "_idx"], fps[name + "_temp"] = xyz, idx.cpu().numpy(), temp.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "fps.npz"), **fps)

    # ---- ball_query / group / three_nn / three_interpolate on one small scene
    b, n, m = 2, 2048, 256
    pts = synth.make_batch(b, n)
    xyz = pts[..., :3].copy()
    temp = torch.full((b, n), 1e10, device=dev)
    fidx = torch.empty((b, m), dtype=torch.int32, device=dev)
    p2.furthest_point_sampling_wrapper(b, n, m, t(xyz), temp, fidx)
    new_xyz = np.take_along_axis(xyz, fidx.cpu().numpy()[..., None].astype(np.int64), 1)
    new_xyz[:, :5] += 300.0  # empty balls
    d = {"xyz": xyz, "new_xyz": new_xyz}
    for r, k in ((0.5, 16), (1.0, 32), (4.0, 8)):
        idx = torch.zeros((b, m, k), dtype=torch.int32, device=dev)
        p2.ball_query_wrapper(b, n, m, r, k, t(new_xyz), t(xyz), idx)
        d[f"bq_{r}_{k}"] = idx.cpu().numpy()
    feat = rng.normal(size=(b, 7, n)).astype(np.float32)
    gidx = t(d["bq_1.0_32"])
    out = torch.empty((b, 7, m, 32), device=dev)
    p2.group_points_wrapper(b, 7, n, m, 32, t(feat), gidx, out)
    d["feat"], d["grouped"] = feat, out.cpu().numpy()
    gout = torch.empty((b, 7, m), device=dev)
    p2.gather_points_wrapper(b, 7, n, m, t(feat), fidx, gout)
    d["fps_idx"], d["gathered"] = fidx.cpu().numpy(), gout.cpu().numpy()
    dist2 = torch.empty((b, n, 3), device=dev)
    nidx = torch.empty((b, n, 3), dtype=torch.int32, device=dev)
    known = np.take_along_axis(xyz, fidx.cpu().numpy()[..., None].astype(np.int64), 1)
    p2.three_nn_wrapper(b, n, m, t(xyz), t(known), dist2, nidx)
    d["known"], d["nn_dist2"], d["nn_idx"] = known, dist2.cpu().numpy(), nidx.cpu().numpy()
    w = rng.uniform(0.05, 1, (b, n, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    kfeat = rng.normal(size=(b, 9, m)).astype(np.float32)
    iout = torch.empty((b, 9, n), device=dev)
    p2.three_interpolate_wrapper(b, 9, m, n, t(kfeat), nidx, t(w), iout)
    d["kfeat"], d["weight"], d["interp"] = kfeat, w, iout.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "pointnet2.npz"), **d)

    # ---- iou3d
    def boxes(nb, spread):
        cx, cy = rng.uniform(-spread, spread, nb), rng.uniform(-spread, spread, nb)
        l, w_ = 3.9 + rng.normal(0, 0.3, nb), 1.6 + rng.normal(0, 0.1, nb)
        ry = rng.uniform(-np.pi, np.pi, nb)
        return np.stack([cx - l / 2, cy - w_ / 2, cx + l / 2, cy + w_ / 2, ry], 1).astype(np.float32)

    a, bb = boxes(150, 6.0), boxes(130, 6.0)
    ov, io = torch.zeros((150, 130), device=dev), torch.zeros((150, 130), device=dev)
    iou.boxes_overlap_bev_gpu(t(a), t(bb), ov)
    iou.boxes_iou_bev_gpu(t(a), t(bb), io)
    d = {"a": a, "b": bb, "overlap": ov.cpu().numpy(), "iou": io.cpu().numpy()}
    nb = 700
    sb = boxes(nb, 7.0)  # already "sorted by score"
    for th in (0.1, 0.5, 0.85):
        keep = torch.zeros(nb, dtype=torch.int64)
        num = iou.nms_gpu(t(sb), keep, th)
        d[f"nms_{th}"] = keep[:num].numpy()
        keep = torch.zeros(nb, dtype=torch.int64)
        num = iou.nms_normal_gpu(t(sb), keep, th)
        d[f"nmsn_{th}"] = keep[:num].numpy()
    d["nms_boxes"] = sb
    np.savez_compressed(os.path.join(out_dir, "iou3d.npz"), **d)

    # ---- roipool3d
    b, n, m, c, s = 2, 2048, 48, 3, 32
    pts = synth.make_batch(b, n, seed=77)
    xyz = pts[..., :3].copy()
    feat = rng.normal(size=(b, n, c)).astype(np.float32)
    bx = np.stack([synth.make_boxes(xyz[i], m, seed=5 + i) for i in range(b)], 0)
    bx[:, :6, 0] += 400.0
    bx[:, -1, 3:6] = 25.0
    pooled = torch.zeros((b, m, s, 3 + c), device=dev)
    flag = torch.zeros((b, m), dtype=torch.int32, device=dev)
    roi.forward(t(xyz), t(bx), t(feat), pooled, flag)
    np.savez_compressed(os.path.join(out_dir, "roipool3d.npz"), xyz=xyz, feat=feat, boxes=bx, pooled=pooled.cpu().numpy(),
                        flag=flag.cpu().numpy())
    torch.cuda.synchronize()
    with open(os.path.join(out_dir, "README.md"), "w") as f:
        f.write("# Golden fixtures\n\nProduced by `tools/make_golden.py` running the reference's own kernels "
                "(oracle/_ref: unmodified /root/reference sources, nvcc 12.9 -O2, sm_100a) on "
                f"{meta['gpu']} with torch {meta['torch']}.\nInputs are seeded; every array named like an op output "
                "is the reference's output for the inputs stored beside it.\n")
    print("golden written to", out_dir, {k: os.path.getsize(os.path.join(out_dir, k)) for k in os.listdir(out_dir)})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
