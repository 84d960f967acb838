"""Workload for compute-sanitizer (race / memory checking evidence):

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize.py

Small shapes (the tools slow kernels down 10-100x) that still reach every kernel family of the forward path: cluster /
DSMEM FPS (flat and two-level) and the bucketed FPS, cell-grid and shared-memory ball query, ring-search three_nn with
the weight epilogue, grouping, interpolation (plain and affine), the tcgen05 layer kernel and the fused SA kernel
(TMEM reuse in place), iou3d / NMS / roipool3d, the f2-f4 kernels, one training forward / backward (training layers, gradient kernels), and a StreamedBackboneRunner whose graphs replay
concurrently on per-buffer scratch arenas.  Outputs are compared with the plain forward so that a silent corruption
would also fail the run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ws3d_b200 import data_utils, iou3d_utils, label_utils, models, native, pointnet2_utils, roipool3d_utils, synth, train_functions  # noqa: E402
from ws3d_b200.graphs import StreamedBackboneRunner  # noqa: E402

dev = "cuda:0"
quick = "--quick" in sys.argv


problems = []


def main():
    torch.manual_seed(0)
    # ---- irregular ops at sizes that select each kernel variant
    for n, m in ((4096, 1024), (16384, 512), (1024, 256)) if not quick else ((4096, 512),):
        pts = torch.from_numpy(synth.make_batch(2, n)).to(dev)
        xyz = pts[..., :3].contiguous()
        feat = pts[..., 3:].transpose(1, 2).contiguous()
        ref_idx = None
        for mode in (0, 1, 2):
            prev = native.set_fps_mode(mode)
            idx, new_xyz = pointnet2_utils.sample_and_gather(xyz, m)
            native.set_fps_mode(prev)
            if ref_idx is None:
                ref_idx = idx.clone()
            elif not torch.equal(idx, ref_idx):
                problems.append(f"FPS mode {mode} differs from mode 0 at n={n} m={m}: {int((idx != ref_idx).sum())} indices")
        i0, i1 = pointnet2_utils.ball_query_pair((0.5, 1.0), (16, 32), xyz, new_xyz)
        g = pointnet2_utils.group_concat(xyz, new_xyz, feat, i1, True)
        nn_idx, w = pointnet2_utils.three_nn_weights(xyz, new_xyz)
        pointnet2_utils.three_interpolate(g[:, :, :, 0].contiguous(), nn_idx, w)
    # ---- throughput sampler with several clouds per CTA (csrc/fps_smem.cu): sub-blocks of 16 / 8 / 4 warps on their own named
    #      barriers, odd batches (a sub-block of the last CTA leaves at once); few samples keep the tool's run time short
    for b, n, m in ((3, 16384, 96), (5, 8192, 96), (9, 2048, 96)):
        xyz = torch.from_numpy(np.ascontiguousarray(synth.make_batch(b, n)[..., :3])).to(dev)
        outs = []
        for mode in (0, 1):
            prev = native.set_fps_mode(mode)
            outs.append(pointnet2_utils.sample_and_gather(xyz, m)[0].clone())
            native.set_fps_mode(prev)
        if not torch.equal(outs[0], outs[1]):
            problems.append(f"throughput FPS differs from the latency kernel at b={b} n={n} m={m}")
    # ---- whole backbone (fused SA kernel, layer kernel, affine gathers) eager, then streamed with 3 batches in flight
    cfg = {"NPOINTS": [512, 128, 32, 8], "RADIUS": models.RPN_SA_CONFIG["RADIUS"], "NSAMPLE": models.RPN_SA_CONFIG["NSAMPLE"],
           "MLPS": [[[8, 8, 16], [8, 8, 16]], [[16, 16, 32], [16, 24, 32]], [[32, 32, 64], [32, 48, 64]], [[64, 64, 96], [64, 64, 96]]]}
    model = models.Pointnet2MSG(input_channels=1, sa_config=cfg, fp_mlps=[[32, 32], [48, 48], [64, 64], [64, 64]]).to(dev).eval()
    batches = [torch.from_numpy(synth.make_batch(2, 4096, first_scene=5 * k)).to(dev) for k in range(4)]
    with torch.no_grad():
        want = [model(b)[1].clone() for b in batches]
        again = [model(b)[1].clone() for b in batches]
        prev = native.set_fps_mode(1)
        plans = [model.coordinate_phase(b) for b in batches]
        native.set_fps_mode(prev)
        two_phase = [model.feature_phase(b, p)[1].clone() for b, p in zip(batches, plans)]
    torch.cuda.synchronize()
    for k in range(4):
        if not torch.equal(want[k], again[k]):
            problems.append(f"plain forward of batch {k} is not reproducible: max diff {float((want[k] - again[k]).abs().max())}")
        if not torch.equal(want[k], two_phase[k]):
            problems.append(f"eager two-phase forward of batch {k} differs from the plain forward: {float((want[k] - two_phase[k]).abs().max())}")
    runner = StreamedBackboneRunner(model, batches[0], lookahead=2, feature_streams=2)
    runner.submit(batches[0])
    runner.submit(batches[1])
    runner.fork()
    for k in range(4):
        got = runner.complete(consume=lambda o: o.clone())
        if k + 2 < 4:
            runner.submit(batches[k + 2])
        runner.join()
        torch.cuda.synchronize()
        if not torch.equal(got, want[k]):
            problems.append(f"streamed batch {k} differs from the plain forward: max diff {float((got - want[k]).abs().max())}, "
                            f"{int((got != want[k]).sum())} of {got.numel()} values")
    # ---- training mode: the tcgen05 GEMM with batch statistics, fused BN + ReLU (+ pool) forward / backward, split-K weight
    #      gradient, the grouping gradient and three_interpolate_grad as a gather over the inverse stencil
    trainee = models.Pointnet2MSG(input_channels=1, sa_config=cfg, fp_mlps=[[32, 32], [48, 48], [64, 64], [64, 64]]).to(dev).train()
    grads = []
    for _ in range(2):
        trainee.zero_grad(set_to_none=True)
        trainee(batches[0])[1].square().mean().backward()
        grads.append(torch.cat([q.grad.flatten() for q in trainee.parameters() if q.grad is not None]).clone())
    torch.cuda.synchronize()
    if not torch.isfinite(grads[0]).all():
        problems.append("training step produced non-finite gradients")
    # ---- Stage-2 widths: the fused kernel with one CTA per SM and four column shares (512 threads), rows packed from
    #      channel-major features, fused levels chained through point-major rows
    from ws3d_b200 import workloads
    from ws3d_b200.pointnet2_modules import sa_stack_forward
    s2 = workloads.Stage2SA().to(dev).eval()
    sx = torch.from_numpy((np.random.default_rng(3).normal(0, 1, (3, 512, 3)) * np.array([1.2, 0.6, 2.2])).astype(np.float32)).to(dev)
    sf = torch.randn(3, 128, 512, device=dev)
    with torch.no_grad():
        a = sa_stack_forward(s2.SA_modules, sx, sf)[1].clone()
        b2 = sa_stack_forward(s2.SA_modules, sx, sf)[1].clone()
    torch.cuda.synchronize()
    if not torch.equal(a, b2):
        problems.append(f"Stage-2 stack is not reproducible: max diff {float((a - b2).abs().max())}")
    rpn = models.RPN().to(dev).eval()
    with torch.no_grad():
        rpn(torch.from_numpy(synth.make_batch(1, 16384 if not quick else 4096)).to(dev))
    # ---- iou3d / roipool3d / next rows
    scene = synth.make_scene(0)
    boxes3d = torch.from_numpy(synth.make_boxes(scene[:, :3], 700)).to(dev)
    scores = torch.rand(700, device=dev)
    from ws3d_b200 import kitti_utils
    bev = kitti_utils.boxes3d_to_bev_torch(boxes3d)
    iou3d_utils.boxes_iou_bev(bev, bev)
    iou3d_utils.nms_gpu(bev, scores, 0.85)
    iou3d_utils.nms_normal_gpu(bev, scores, 0.8)
    p = torch.from_numpy(scene[None, :, :3].copy()).to(dev)
    f = torch.from_numpy(scene[None, :, 3:].copy()).to(dev)
    roipool3d_utils.roipool3d_gpu(p, f, boxes3d[None, :128].contiguous(), 1.0, sampled_pt_num=128)
    kitti_utils.boxes3d_to_corners3d_torch(boxes3d)
    pred = (boxes3d + 0.1).requires_grad_(True)
    train_functions.corner_loss(pred, boxes3d).backward()
    gt, cnt = synth.make_gt_boxes(2)
    label_utils.generate_gaussian_training_labels(torch.from_numpy(synth.make_batch(2, 16384)[..., :3].copy()).to(dev),
                                                  torch.from_numpy(gt).to(dev), torch.from_numpy(cnt).to(dev))
    depth = torch.from_numpy(np.abs(scene[:, 2]).astype(np.float32)).to(dev)
    n_near = int((depth < 40).sum())
    perm, order = data_utils.draw_subsample(np.random.RandomState(0), 16384, n_near, 8192)
    _, _, status = data_utils.subsample_points(torch.from_numpy(scene).to(dev), depth, 8192, perm, order, n_near)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    if problems:
        print("sanitize workload: RESULT MISMATCHES under the tool:")
        for p in problems:
            print("  -", p)
        sys.exit(3)
    print("sanitize workload ok")


if __name__ == "__main__":
    main()
