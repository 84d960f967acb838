"""Module-level GPU parity: the B200 SA / FP modules and the whole MSG backbone vs the CPU
composition of oracle ops with the same weights (fp32, TF32 disabled; tolerance 2e-4 absolute on
O(1) activations -- index ops are exact, the difference is cuDNN vs host GEMM summation order)."""
import copy

import numpy as np
import pytest
import torch

from oracle import cpu_backbone

pytestmark = pytest.mark.gpu
dev = "cuda:0"


@pytest.fixture(autouse=True)
def _fp32_math():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _randomize_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def test_backbone_forward_matches_cpu_composition():
    from ws3d_b200 import models, synth
    torch.manual_seed(0)
    cfg = {"NPOINTS": [512, 128, 32, 8], "RADIUS": models.RPN_SA_CONFIG["RADIUS"], "NSAMPLE": models.RPN_SA_CONFIG["NSAMPLE"],
           "MLPS": [[[8, 8, 16], [8, 8, 16]], [[16, 16, 32], [16, 24, 32]], [[32, 32, 64], [32, 48, 64]], [[64, 64, 96], [64, 64, 96]]]}
    fp = [[32, 32], [48, 48], [64, 64], [64, 64]]
    cpu_model = models.Pointnet2MSG(input_channels=1, sa_config=cfg, fp_mlps=fp).eval()
    _randomize_bn(cpu_model, 1)
    gpu_model = copy.deepcopy(cpu_model).to(dev).eval()
    pts = synth.make_batch(2, 2048)
    with torch.no_grad():
        gx, gf = gpu_model(torch.from_numpy(pts).to(dev))
    cx, cf = cpu_backbone.backbone_forward(cpu_model, pts)
    np.testing.assert_array_equal(gx.cpu().numpy(), cx)
    np.testing.assert_allclose(gf.cpu().numpy(), cf, rtol=1e-3, atol=2e-4)


def test_sa_module_backward_runs_and_matches_unfused():
    """Gradients through the fused SA layer equal those of the reference-style unfused sequence."""
    from ws3d_b200 import pointnet2_modules, pointnet2_utils, synth
    torch.manual_seed(1)
    sa = pointnet2_modules.PointnetSAModuleMSG(npoint=256, radii=[0.5, 1.0], nsamples=[16, 32],
                                              mlps=[[4, 8, 16], [4, 8, 16]], use_xyz=True).to(dev).train()
    pts = torch.from_numpy(synth.make_batch(2, 4096)).to(dev)
    xyz = pts[..., :3].contiguous()
    feat = torch.randn(2, 4, 4096, device=dev, requires_grad=True)
    new_xyz, out = sa(xyz, feat)
    out.square().mean().backward()
    g_fused = feat.grad.clone()
    # unfused: FPS -> gather -> per-scale ball_query / group / subtract / cat (pointnet2_modules.py:30-55)
    feat2 = feat.detach().clone().requires_grad_(True)
    idx = pointnet2_utils.furthest_point_sample(xyz, 256)
    nx = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    assert torch.equal(nx, new_xyz)
    outs = []
    for g, mlp in zip(sa.groupers, sa.mlps):
        bi = pointnet2_utils.ball_query(g.radius, g.nsample, xyz, nx)
        gx = pointnet2_utils.grouping_operation(xyz.transpose(1, 2).contiguous(), bi) - nx.transpose(1, 2).unsqueeze(-1)
        gf = pointnet2_utils.grouping_operation(feat2, bi)
        y = mlp(torch.cat([gx, gf], dim=1))
        outs.append(torch.nn.functional.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1))
    out2 = torch.cat(outs, dim=1)
    torch.testing.assert_close(out2, out, rtol=1e-4, atol=1e-5)
    out2.square().mean().backward()
    torch.testing.assert_close(feat2.grad, g_fused, rtol=1e-3, atol=1e-6)


def test_dropin_modules_expose_reference_tables():
    import ws3d_b200
    ws3d_b200.install_dropins()
    import iou3d_cuda
    import pointnet2_cuda
    import roipool3d_cuda
    for name in ("ball_query_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "gather_points_wrapper",
                 "gather_points_grad_wrapper", "furthest_point_sampling_wrapper", "three_nn_wrapper",
                 "three_interpolate_wrapper", "three_interpolate_grad_wrapper"):
        assert callable(getattr(pointnet2_cuda, name))
    for name in ("boxes_overlap_bev_gpu", "boxes_iou_bev_gpu", "nms_gpu", "nms_normal_gpu"):
        assert callable(getattr(iou3d_cuda, name))
    for name in ("forward", "forward_slow", "pts_in_boxes3d_cpu", "roipool3d_cpu"):
        assert callable(getattr(roipool3d_cuda, name))
    # legacy-constructor call pattern of the reference wrapper (pointnet2_utils.py:24-29)
    xyz = torch.rand(2, 512, 3, device=dev)
    out = torch.cuda.IntTensor(2, 64)
    temp = torch.cuda.FloatTensor(2, 512).fill_(1e10)
    assert pointnet2_cuda.furthest_point_sampling_wrapper(2, 512, 64, xyz, temp, out) == 1
    assert int(out[:, 0].abs().sum()) == 0


def test_cuda_graph_replay_equals_eager_two_stream_forward():
    """ws3d_b200.graphs.CudaGraphRunner: the captured two-stream forward replays to the same features as the eager
    call, for a new input copied into the static buffer, and the single-stream schedule gives the same result."""
    import os
    from ws3d_b200 import models, synth
    from ws3d_b200.graphs import CudaGraphRunner
    torch.manual_seed(0)
    model = models.Pointnet2MSG(input_channels=1).to(dev).eval()
    a = torch.from_numpy(synth.make_batch(2, 4096)).to(dev)
    b = torch.from_numpy(synth.make_batch(2, 4096, first_scene=7)).to(dev)
    with torch.no_grad():
        want_a, want_b = model(a)[1].clone(), model(b)[1].clone()
        os.environ["WS3D_TWO_STREAMS"] = "0"
        try:
            single = model(b)[1].clone()
        finally:
            del os.environ["WS3D_TWO_STREAMS"]
    assert torch.equal(single, want_b)
    runner = CudaGraphRunner(lambda x: model(x)[1], a)
    assert torch.equal(runner(a), want_a)
    assert torch.equal(runner(b), want_b)
    assert torch.equal(runner(a), want_a)


def test_pipelined_runner_equals_unpipelined_forward():
    """graphs.PipelinedBackboneRunner (level-1 FPS of batch i+1 beside the rest of batch i, two alternating CUDA
    graphs) returns, for every batch of a stream of batches, exactly what the plain forward returns."""
    from ws3d_b200 import models, synth
    from ws3d_b200.graphs import PipelinedBackboneRunner
    torch.manual_seed(0)
    cfg = {"NPOINTS": [512, 128, 32, 8], "RADIUS": models.RPN_SA_CONFIG["RADIUS"], "NSAMPLE": models.RPN_SA_CONFIG["NSAMPLE"],
           "MLPS": [[[8, 8, 16], [8, 8, 16]], [[16, 16, 32], [16, 24, 32]], [[32, 32, 64], [32, 48, 64]], [[64, 64, 96], [64, 64, 96]]]}
    fp = [[32, 32], [48, 48], [64, 64], [64, 64]]
    model = models.Pointnet2MSG(input_channels=1, sa_config=cfg, fp_mlps=fp).to(dev).eval()
    _randomize_bn(model, 2)
    batches = [torch.from_numpy(synth.make_batch(2, 2048, first_scene=10 * k)) for k in range(5)]
    with torch.no_grad():
        want = [model(b.to(dev))[1].clone() for b in batches]
    runner = PipelinedBackboneRunner(model, batches[0].to(dev))
    pinned = [b.pin_memory() for b in batches]
    runner.prefetch(pinned[0])                       # host batches: H2D on the runner's copy stream
    got = []
    for k in range(len(batches)):
        out = runner.step(pinned[k + 1] if k + 1 < len(batches) else None)
        got.append(out.clone())
    torch.cuda.synchronize()
    for k, (g, w) in enumerate(zip(got, want)):
        assert torch.equal(g, w), f"batch {k} differs"
    # a second pass over device-resident batches through stage_next(), issued while the previous step is in flight
    runner.prefetch(batches[4].to(dev))
    runner.stage_next(batches[3].to(dev))
    a = runner.step().clone()
    b = runner.step().clone()
    torch.cuda.synchronize()
    assert torch.equal(a, want[4]) and torch.equal(b, want[3])


def test_streamed_runner_equals_plain_forward():
    """graphs.StreamedBackboneRunner (coordinate phase -- FPS in throughput mode, ball queries, stencils -- two batches
    ahead on its own streams, feature phase on the caller's) returns exactly the plain forward's output for every
    batch of a stream of batches."""
    from ws3d_b200 import models, native, synth
    from ws3d_b200.graphs import StreamedBackboneRunner
    torch.manual_seed(0)
    cfg = {"NPOINTS": [512, 128, 32, 8], "RADIUS": models.RPN_SA_CONFIG["RADIUS"], "NSAMPLE": models.RPN_SA_CONFIG["NSAMPLE"],
           "MLPS": [[[8, 8, 16], [8, 8, 16]], [[16, 16, 32], [16, 24, 32]], [[32, 32, 64], [32, 48, 64]], [[64, 64, 96], [64, 64, 96]]]}
    fp = [[32, 32], [48, 48], [64, 64], [64, 64]]
    model = models.Pointnet2MSG(input_channels=1, sa_config=cfg, fp_mlps=fp).to(dev).eval()
    _randomize_bn(model, 3)
    batches = [torch.from_numpy(synth.make_batch(2, 4096, first_scene=7 * k)) for k in range(6)]
    with torch.no_grad():
        want = [model(b.to(dev))[1].clone() for b in batches]
    runner = StreamedBackboneRunner(model, batches[0].to(dev), lookahead=2)
    pinned = [b.pin_memory() for b in batches]
    runner.submit(pinned[0])
    runner.submit(pinned[1])
    got = []
    for k in range(len(batches)):
        got.append(runner.complete().clone())
        if k + 2 < len(batches):
            runner.submit(pinned[k + 2])
    torch.cuda.synchronize()
    for k, (g, w) in enumerate(zip(got, want)):
        assert torch.equal(g, w), f"batch {k} differs"
    assert native.set_fps_mode(0) == 0          # the runner restores the calling thread's mode
    # feature phases of consecutive batches on two internal streams; results consumed on those streams
    runner2 = StreamedBackboneRunner(model, batches[0].to(dev), lookahead=3, feature_streams=2)
    for k in range(3):
        runner2.submit(pinned[k])
    runner2.fork()
    got2 = []
    for k in range(len(batches)):
        got2.append(runner2.complete(consume=lambda o: o.clone()))
        if k + 3 < len(batches):
            runner2.submit(pinned[k + 3])
    runner2.join()
    torch.cuda.synchronize()
    for k, (g, w) in enumerate(zip(got2, want)):
        assert torch.equal(g, w), f"batch {k} differs (two feature streams)"
    # cold start: a batch submitted into an (almost) empty pipeline replays the latency-sampler capture of its coordinate
    # phase; here the pipeline runs dry after every other batch, so both captures of every buffer set are used
    for cold in (0, 1, 3):
        runner3 = StreamedBackboneRunner(model, batches[0].to(dev), lookahead=2, feature_streams=2, cold_start=cold)
        got3 = []
        runner3.fork()
        for k in range(0, len(batches), 2):
            runner3.submit(pinned[k])
            runner3.submit(pinned[k + 1])
            got3.append(runner3.complete(consume=lambda o: o.clone()))
            got3.append(runner3.complete(consume=lambda o: o.clone()))
        runner3.join()
        torch.cuda.synchronize()
        for k, (g, w) in enumerate(zip(got3, want)):
            assert torch.equal(g, w), f"batch {k} differs (cold_start={cold})"


def test_fps_modes_are_bit_identical():
    from ws3d_b200 import native, pointnet2_utils, synth
    xyz = torch.from_numpy(np.ascontiguousarray(synth.make_batch(3, 8192)[..., :3])).to(dev)
    outs = []
    for mode in (0, 1, 2):
        prev = native.set_fps_mode(mode)
        try:
            idx, nx = pointnet2_utils.sample_and_gather(xyz, 2048)
        finally:
            native.set_fps_mode(prev)
        outs.append((idx.clone(), nx.clone()))
    for idx, nx in outs[1:]:
        assert torch.equal(idx, outs[0][0]) and torch.equal(nx, outs[0][1])


def test_streamed_runner_at_bench_config_equals_plain_forward():
    """The HEADLINE configuration of bench.py -- full weaklyRPN.yaml network, 16 clouds x 16384 points, 7 batches in
    flight (lookahead 6), two feature streams, persistent kernels capped at 100 SMs, TF32 MLPs, the first two batches on
    the latency samplers (cold start) -- returns exactly the plain forward's tensor for every batch of a stream of 14
    batches (two alternating inputs, host and resident)."""
    from ws3d_b200 import models, native, synth
    from ws3d_b200.graphs import StreamedBackboneRunner
    torch.backends.cudnn.allow_tf32 = True           # (the autouse fixture restores the previous value)
    torch.manual_seed(0)
    model = models.Pointnet2MSG(input_channels=1).to(dev).eval()
    _randomize_bn(model, 11)
    hosts = [torch.from_numpy(synth.make_batch(16, 16384, first_scene=16 * k)).pin_memory() for k in range(2)]
    with torch.no_grad():
        want = [model(h.to(dev))[1].clone() for h in hosts]
    prev = native.set_sm_budget(100)
    try:
        look = 6
        runner = StreamedBackboneRunner(model, hosts[0].to(dev), lookahead=look, feature_streams=2, cold_start=2)
        n = 14
        for j in range(look):
            runner.submit(hosts[j % 2])
        runner.fork()
        bad = []
        for j in range(n):
            ok = runner.complete(consume=lambda o, j=j: torch.equal(o, want[j % 2]))
            if not ok:
                bad.append(j)
            if j + look < n:
                runner.submit(hosts[(j + look) % 2] if j % 3 else hosts[(j + look) % 2].to(dev))
        runner.join()
        torch.cuda.synchronize()
        assert not bad, f"batches {bad} differ from the plain forward"
    finally:
        native.set_sm_budget(prev)


def test_full_config_fp32_forward_matches_cpu_oracle_composition():
    """The full-size network (weaklyRPN.yaml shapes) on one 16384-point cloud, FP32 MLPs, against the CPU composition of
    oracle ops: sampled coordinates exact, features allclose(2e-4)."""
    from ws3d_b200 import models, synth
    torch.manual_seed(0)
    cpu_model = models.Pointnet2MSG(input_channels=1).eval()
    _randomize_bn(cpu_model, 2)
    gpu_model = copy.deepcopy(cpu_model).to(dev).eval()
    pts = synth.make_batch(1, 16384, first_scene=40)
    with torch.no_grad():
        gx, gf = gpu_model(torch.from_numpy(pts).to(dev))
    cx, cf = cpu_backbone.backbone_forward(cpu_model, pts)
    np.testing.assert_array_equal(gx.cpu().numpy(), cx)
    scale = float(np.abs(cf).max())
    np.testing.assert_allclose(gf.cpu().numpy(), cf, rtol=2e-4, atol=2e-4 * scale)


def test_three_nn_weights_equals_torch_formula():
    """ws3d_three_nn_weights: idx equal to three_nn's, weights equal to the five torch kernels of
    pointnet2_modules.py:139-144 (every step is one IEEE float32 operation; torch's 3-term sum may associate
    differently, hence one ulp of slack)."""
    from ws3d_b200 import pointnet2_utils, synth
    for n, m in ((16384, 4096), (1024, 256), (256, 64), (300, 2)):
        pts = torch.from_numpy(synth.make_batch(2, n, first_scene=3)).to(dev)
        unknown = pts[..., :3].contiguous()
        known = unknown[:, torch.randperm(n, generator=torch.Generator().manual_seed(n))[:m].to(dev)].contiguous()
        dist, idx = pointnet2_utils.three_nn(unknown, known)
        recip = 1.0 / (dist + 1e-8)
        want = recip / torch.sum(recip, dim=2, keepdim=True)
        idx2, weight = pointnet2_utils.three_nn_weights(unknown, known)
        assert torch.equal(idx, idx2)
        torch.testing.assert_close(weight, want, rtol=3e-7, atol=1e-9)
        exact = float((weight == want).float().mean())
        assert exact > 0.5, exact     # the same arithmetic up to the association of the 3-term sum


def test_ball_query_writes_every_row_of_uninitialised_output():
    """Rows of centres without any neighbour are zero-filled by the kernels (the reference relies on the caller's
    zero fill): every path -- cell grid, shared-memory scan, generic -- on garbage-filled idx buffers."""
    import oracle
    from ws3d_b200 import native, synth
    for n, m in ((16384, 512), (1024, 128), (20000, 64)):
        pts = synth.make_batch(2, n, first_scene=9)
        xyz = np.ascontiguousarray(pts[..., :3])
        new_xyz = xyz[:, :m].copy()
        new_xyz[:, ::3] += 500.0                       # every third centre is far from every point
        new_xyz[0, 1] = np.nan
        tx, tn = torch.from_numpy(xyz).to(dev), torch.from_numpy(new_xyz).to(dev)
        idx = torch.full((2, m, 16), -7, dtype=torch.int32, device=dev)
        native.ball_query_wrapper(2, n, m, 0.5, 16, tn, tx, idx)
        np.testing.assert_array_equal(idx.cpu().numpy(), oracle.ball_query(0.5, 16, xyz, new_xyz))
        i0 = torch.full((2, m, 16), -7, dtype=torch.int32, device=dev)
        i1 = torch.full((2, m, 32), -7, dtype=torch.int32, device=dev)
        native.ball_query2(2, n, m, 0.5, 16, 1.0, 32, tn, tx, i0, i1)
        np.testing.assert_array_equal(i0.cpu().numpy(), oracle.ball_query(0.5, 16, xyz, new_xyz))
        np.testing.assert_array_equal(i1.cpu().numpy(), oracle.ball_query(1.0, 32, xyz, new_xyz))
        assert int((i0[:, ::3] != 0).sum()) == 0


def test_rpn_fused_heads_equal_separate_heads():
    """models.RPN in inference: both heads as one chain of two launches (stacked / block-diagonal weights) against the
    per-head tensor-core path and the PyTorch modules."""
    from ws3d_b200 import fused_mlp, models
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(3)
    rpn = models.RPN().to(dev).eval()
    _randomize_bn(rpn, 4)
    feats = torch.randn(2, 128, 4096, device=dev)
    with torch.no_grad():
        both = fused_mlp.FoldedHeads([rpn.rpn_cls_layer, rpn.rpn_reg_layer])(feats)
        cls_sep = rpn._head("rpn_cls_layer", feats)
        reg_sep = rpn._head("rpn_reg_layer", feats)
        cls_pt, reg_pt = rpn.rpn_cls_layer(feats), rpn.rpn_reg_layer(feats)
    torch.testing.assert_close(both[:, :1], cls_sep, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(both[:, 1:], reg_sep, rtol=1e-6, atol=1e-6)
    assert float((both[:, :1] - cls_pt).abs().max()) < 3e-3 * float(cls_pt.abs().max())
    assert float((both[:, 1:] - reg_pt).abs().max()) < 3e-3 * max(1e-3, float(reg_pt.abs().max()))


def test_scratch_release_frees_retired_buffers():
    from ws3d_b200 import native, pointnet2_utils, synth
    torch.cuda.synchronize()
    native.release_scratch(everything=True)
    assert native.scratch_bytes() == 0
    for n in (4096, 16384):      # the second call outgrows the first call's cell-grid scratch
        xyz = torch.from_numpy(np.ascontiguousarray(synth.make_batch(4, n)[..., :3])).to(dev)
        pointnet2_utils.ball_query(0.5, 16, xyz, xyz[:, :256].contiguous())
    assert native.scratch_bytes(retired_only=True) > 0
    live = native.scratch_bytes() - native.scratch_bytes(retired_only=True)
    native.release_scratch()
    assert native.scratch_bytes(retired_only=True) == 0 and native.scratch_bytes() == live
    native.release_scratch(everything=True)
    assert native.scratch_bytes() == 0
    xyz = torch.from_numpy(np.ascontiguousarray(synth.make_batch(1, 4096)[..., :3])).to(dev)
    pointnet2_utils.ball_query(0.5, 16, xyz, xyz[:, :64].contiguous())       # the library re-allocates on demand
    torch.cuda.synchronize()


def test_fused_sa_point_major_rows_path_is_bit_identical():
    """The fused SA kernels gathering from point-major rows (the (B,N,4) cloud itself at level 1, the previous level's
    point-major output at level 2) give exactly the channel-major gather's result, and the emitted rows are
    [centre xyz | pooled channels of both scales | zeros]."""
    import os

    from ws3d_b200 import models, synth
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(0)
    model = models.Pointnet2MSG(input_channels=1).to(dev).eval()
    _randomize_bn(model, 6)
    pts = torch.from_numpy(synth.make_batch(3, 16384, first_scene=21)).to(dev)
    with torch.no_grad():
        os.environ["WS3D_SA_ROWS"] = "0"
        try:
            want = model(pts)[1].clone()
        finally:
            os.environ.pop("WS3D_SA_ROWS", None)
        got = model(pts)[1]
        assert torch.equal(got, want)
        xyz, feat = model._break_up_pc(pts)
        sa1 = model.SA_modules[0]
        nx, nf, rows = sa1(xyz, feat, rows=pts, want_rows=True)
        assert rows is not None and rows.shape == (3, 4096, 104)
        assert torch.equal(rows[..., :3], nx)
        assert torch.equal(rows[..., 3:99], nf.transpose(1, 2))
        assert float(rows[..., 99:].abs().max()) == 0.0
        nx2, nf2 = model.SA_modules[1](nx, nf, rows=rows)
        nx2b, nf2b = model.SA_modules[1](nx, nf)
        assert torch.equal(nx2, nx2b) and torch.equal(nf2, nf2b)
        # a ragged shape: centres per cloud not a multiple of the tile, cloud rows wider than the operand
        pts5 = torch.cat([pts[:, :5000], torch.zeros(3, 5000, 4, device=dev)], dim=2).contiguous()      # (3, 5000, 8): ld = 8 > 3 + 1
        sa = models.Pointnet2MSG(input_channels=1).SA_modules[0].to(dev).eval()
        sa.npoint = 1000
        x5, f5 = pts5[..., :3].contiguous(), pts5[..., 3:4].transpose(1, 2).contiguous()
        a = sa(x5, f5)
        b = sa(x5, f5, rows=pts5, want_rows=True)
        assert torch.equal(a[1], b[1]) and torch.equal(b[2][..., 3:99], b[1].transpose(1, 2))


def test_pack_rows_and_stage2_stack_chain_equal_module_by_module():
    """pack_rows = [xyz | features^T | zeros] exactly; the Stage-2 stack through sa_stack_forward (rows packed once, fused
    levels hand point-major rows to each other) returns what the modules return one after the other on channel-major
    features (WS3D_SA_ROWS=0), bit for bit: the gather source changes, not the values."""
    import os

    from ws3d_b200 import pointnet2_utils, workloads
    from ws3d_b200.pointnet2_modules import sa_stack_forward
    torch.manual_seed(0)
    B, N, C = 24, 512, 128
    rng = np.random.default_rng(5)
    xyz = torch.from_numpy((rng.normal(0, 1, (B, N, 3)) * np.array([1.2, 0.6, 2.2])).astype(np.float32)).to(dev)
    feats = torch.randn(B, C, N, device=dev)
    rows = pointnet2_utils.pack_rows(xyz, feats)
    assert rows.shape == (B, N, 136)
    want = torch.cat([xyz, feats.transpose(1, 2), torch.zeros(B, N, 136 - 3 - C, device=dev)], dim=2)
    assert torch.equal(rows, want)
    odd = pointnet2_utils.pack_rows(xyz[:, :77].contiguous(), feats[:, :5, :77].contiguous())       # ragged tiles
    assert torch.equal(odd, torch.cat([xyz[:, :77], feats[:, :5, :77].transpose(1, 2)], dim=2))
    assert torch.equal(pointnet2_utils.pack_rows(xyz, None)[..., :3], xyz)
    model = workloads.Stage2SA().to(dev).eval()
    with torch.no_grad():
        got = sa_stack_forward(model.SA_modules, xyz, feats)[1]
        os.environ["WS3D_SA_ROWS"] = "0"
        try:
            x, f = xyz, feats
            for sa in model.SA_modules:
                x, f = sa(x, f)
        finally:
            os.environ.pop("WS3D_SA_ROWS", None)
    assert torch.equal(got, f)
