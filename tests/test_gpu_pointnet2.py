"""GPU parity tests for the pointnet2 ops: CUDA path (through the C ABI) vs the CPU oracle and,
when oracle/_ref was built, vs the reference's own kernels recompiled for sm_100a.
Bar: bit-exact for every index-producing op; <= 1e-5 (in practice exact) for gathered /
interpolated values; atomics-based gradients <= 1e-5 relative."""
import numpy as np
import pytest
import torch

import oracle
from refmods import require_ref

pytestmark = pytest.mark.gpu

dev = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _cloud(rng, b, n, kind="uniform"):
    if kind == "uniform":
        return rng.uniform(-10, 10, (b, n, 3)).astype(np.float32)
    if kind == "grid":  # many exact ties and duplicate points
        return rng.integers(-3, 4, (b, n, 3)).astype(np.float32)
    if kind == "scene":
        from ws3d_b200 import synth
        return synth.make_batch(b, n)[..., :3].copy()
    raise ValueError(kind)


FPS_CASES = [
    (1, 16384, 4096, "scene"), (3, 4096, 1024, "uniform"), (2, 1024, 256, "uniform"), (2, 256, 64, "uniform"),
    (4, 512, 256, "uniform"), (1, 100, 100, "uniform"), (2, 1000, 333, "uniform"), (1, 8, 8, "uniform"),
    (1, 1, 1, "uniform"), (2, 5000, 100, "uniform"), (2, 2048, 512, "grid"), (1, 300, 300, "grid"),
    (1, 64, 100, "uniform"),  # m > n: the sampler keeps re-selecting (all distances 0)
    (200, 512, 128, "uniform"), (1, 20000, 64, "uniform"), (20, 16384, 16, "uniform"),
    # batches that leave < 2 SMs per cloud take the spatially bucketed single-CTA kernel (csrc/fps_bucket.cu)
    (80, 2048, 128, "uniform"), (76, 4096, 64, "grid"), (80, 5000, 300, "scene"), (76, 16384, 256, "scene"),
    # clusters of 2 / 4 CTAs with the one-level exchange (fps_flat_kernel): n > 4096
    (40, 5000, 300, "scene"), (3, 8192, 2048, "uniform"), (2, 6000, 700, "grid"), (30, 16384, 128, "scene"),
    # automatic dispatch to the paired-sample shared-memory kernel: large clouds with fewer than 4 SMs each, small ones beyond the SM count
    (40, 16384, 96, "scene"), (150, 2048, 64, "uniform"),
]


@pytest.mark.parametrize("b,n,m,kind", FPS_CASES)
def test_fps_matches_oracle_and_reference(b, n, m, kind):
    from ws3d_b200 import native, pointnet2_utils
    rng = np.random.default_rng(b * 1000003 + n * 101 + m)
    xyz = _cloud(rng, b, n, kind)
    exp_idx, exp_temp = oracle.furthest_point_sample(xyz, m, return_temp=True)
    x = _t(xyz)
    idx = pointnet2_utils.furthest_point_sample(x, m)
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (b, m)
    np.testing.assert_array_equal(idx.cpu().numpy(), exp_idx)
    # raw entry point: temp is read and written back like the reference's scratch tensor
    temp = torch.full((b, n), 1e10, device=dev)
    idx2 = torch.empty((b, m), dtype=torch.int32, device=dev)
    native.furthest_point_sampling_wrapper(b, n, m, x, temp, idx2)
    np.testing.assert_array_equal(idx2.cpu().numpy(), exp_idx)
    np.testing.assert_array_equal(temp.cpu().numpy(), exp_temp)
    # fused sampler: coordinates of the selected points
    idx3, new_xyz = pointnet2_utils.sample_and_gather(x, m)
    np.testing.assert_array_equal(idx3.cpu().numpy(), exp_idx)
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), np.take_along_axis(xyz, exp_idx[..., None].astype(np.int64), 1))
    ref = require_ref("pointnet2_cuda")
    if ref is not None:
        rtemp = torch.full((b, n), 1e10, device=dev)
        ridx = torch.empty((b, m), dtype=torch.int32, device=dev)
        ref.furthest_point_sampling_wrapper(b, n, m, x, rtemp, ridx)
        assert torch.equal(ridx, idx)
        assert torch.equal(rtemp, temp)


TILE_CASES = [(2, 16384, 4096, "scene"), (3, 4096, 1024, "uniform"), (2, 2048, 512, "grid"), (2, 5000, 700, "grid"),
              (1, 16384, 300, "uniform"), (3, 8192, 2048, "scene"), (2, 3000, 3000, "uniform"), (1, 2500, 64, "special"),
              # several clouds per CTA (csrc/fps_smem.cu): odd batch at two clouds per CTA, a tail CTA at eight and at four,
              # lattice ties with every bucket full, a batch that fills the sub-blocks of many CTAs
              (3, 16384, 512, "scene"), (17, 4096, 256, "uniform"), (5, 8192, 300, "grid"), (9, 2048, 2048, "grid"),
              (40, 4096, 128, "scene"), (2, 16384, 700, "special"),
              (1, 2048, 2100, "grid")]   # m > n: every distance reaches 0 and the tie key alone decides


@pytest.mark.parametrize("b,n,m,kind", TILE_CASES)
def test_fps_throughput_mode_matches_oracle(b, n, m, kind):
    """ws3d_set_fps_mode(1): the throughput sampler -- Morton buckets with exact culling, running distances in shared
    memory, several clouds per CTA (csrc/fps_smem.cu; WS3D_FPS_SMEM=0: the one-SM-per-cloud kernel of csrc/fps_bucket.cu) --
    against the oracle: indices, coordinates and the written-back temp, incl. duplicates / lattice ties, non-finite
    points, far outliers and a cloud collapsed to a line."""
    from ws3d_b200 import native
    rng = np.random.default_rng(b * 7919 + n + m)
    if kind == "special":
        xyz = rng.uniform(-5, 5, (b, n, 3)).astype(np.float32)
        xyz[:, 100:400, 1:] = 0.0                    # a 1-D segment
        xyz[:, 400:500] = xyz[:, 399:400]            # 100 copies of one point
        xyz[:, 500] = [1e6, -1e6, 1e6]               # far outlier (stretches the Morton box)
        xyz[:, 600, 0] = np.nan
        xyz[:, 601, 2] = np.inf
    else:
        xyz = _cloud(rng, b, n, kind)
    exp_idx, exp_temp = oracle.furthest_point_sample(xyz, m, return_temp=True)
    x = _t(xyz)
    temp = torch.full((b, n), 1e10, device=dev)
    idx = torch.empty((b, m), dtype=torch.int32, device=dev)
    new_xyz = torch.empty((b, m, 3), device=dev)
    prev = native.set_fps_mode(1)
    try:
        native.furthest_point_sampling_gather(b, n, m, x, temp, idx, new_xyz)
    finally:
        native.set_fps_mode(prev)
    np.testing.assert_array_equal(idx.cpu().numpy(), exp_idx)
    np.testing.assert_array_equal(temp.cpu().numpy(), exp_temp)
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), np.take_along_axis(xyz, exp_idx[..., None].astype(np.int64), 1))
    ref = require_ref("pointnet2_cuda")
    if ref is not None and kind != "special":        # (the reference kernel's own treatment of NaN / inf is pinned by the oracle)
        rtemp = torch.full((b, n), 1e10, device=dev)
        ridx = torch.empty((b, m), dtype=torch.int32, device=dev)
        ref.furthest_point_sampling_wrapper(b, n, m, x, rtemp, ridx)
        assert torch.equal(ridx, idx)
        assert torch.equal(rtemp, temp)


def test_fps_throughput_mode_at_bench_size_equals_latency_mode_and_reference():
    """The bench's level-1 sampling call (16 clouds x 16384 points -> 4096 samples) in throughput mode (two clouds per CTA),
    in latency mode and on the reference's own kernel: same indices, same written-back distances; size-independent properties:
    samples distinct, the selection distances never increase."""
    from ws3d_b200 import native, synth
    b, n, m = 16, 16384, 4096
    x = _t(synth.make_batch(b, n)[..., :3].copy())
    outs = []
    for mode in (1, 2):
        temp = torch.full((b, n), 1e10, device=dev)
        idx = torch.empty((b, m), dtype=torch.int32, device=dev)
        nx = torch.empty((b, m, 3), device=dev)
        prev = native.set_fps_mode(mode)
        try:
            native.furthest_point_sampling_gather(b, n, m, x, temp, idx, nx)
        finally:
            native.set_fps_mode(prev)
        outs.append((idx, temp, nx))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    ref = require_ref("pointnet2_cuda")
    if ref is not None:
        rtemp = torch.full((b, n), 1e10, device=dev)
        ridx = torch.empty((b, m), dtype=torch.int32, device=dev)
        ref.furthest_point_sampling_wrapper(b, n, m, x, rtemp, ridx)
        assert torch.equal(ridx, outs[0][0]) and torch.equal(rtemp, outs[0][1])
    idx, _, nx = outs[0]
    srt = torch.sort(idx.long(), dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                      # distinct samples (the scenes have no duplicate points)
    # distance of sample k to the nearest earlier sample = its running distance when it was selected: non-increasing in k
    sel = nx[0].double()
    d = torch.cdist(sel[:512], sel[:512]).square()
    tri = torch.tril(torch.ones_like(d, dtype=torch.bool), diagonal=-1)
    near = torch.where(tri, d, torch.full_like(d, float("inf"))).min(dim=1).values[1:]
    assert bool((near[1:] <= near[:-1] * (1 + 1e-6)).all())


BQ_CASES = [
    (1, 16384, 4096, 0.8, 32, "scene"), (2, 4096, 1024, 0.5, 16, "scene"), (2, 1024, 256, 1.0, 16, "uniform"),
    (3, 1001, 77, 3.0, 32, "uniform"),   # n*12 not 16-byte aligned for b > 0: exercises the non-TMA staging
    (2, 256, 64, 4.0, 32, "uniform"), (1, 50, 7, 100.0, 64, "uniform"), (2, 512, 256, 0.2, 16, "grid"),
    (1, 16384, 64, 0.1, 16, "scene"), (1, 20000, 100, 2.0, 32, "uniform"),  # > smem capacity: generic path
    (1, 3, 5, 1.0, 4, "uniform"),
]


@pytest.mark.parametrize("b,n,m,r,k,kind", BQ_CASES)
def test_ball_query_matches_oracle_and_reference(b, n, m, r, k, kind):
    from ws3d_b200 import pointnet2_utils
    rng = np.random.default_rng(n * 7 + m)
    xyz = _cloud(rng, b, n, kind)
    pick = rng.integers(0, n, (b, m))
    new_xyz = np.take_along_axis(xyz, pick[..., None], 1).copy()
    new_xyz[:, : max(1, m // 8)] += 1000.0  # some centres with no neighbour at all: rows stay zero
    exp = oracle.ball_query(r, k, xyz, new_xyz)
    x, q = _t(xyz), _t(new_xyz)
    idx = pointnet2_utils.ball_query(r, k, x, q)
    np.testing.assert_array_equal(idx.cpu().numpy(), exp)
    ref = require_ref("pointnet2_cuda")
    if ref is not None:
        ridx = torch.zeros((b, m, k), dtype=torch.int32, device=dev)
        ref.ball_query_wrapper(b, n, m, r, k, q, x, ridx)
        assert torch.equal(ridx, idx)


GRID_CASES = ["dense_cluster", "duplicates", "flat", "line", "nonfinite", "tiny_radius", "huge_radius", "offset", "integer_lattice"]


@pytest.mark.parametrize("kind", GRID_CASES)
def test_ball_query_cell_grid_edge_cases(kind):
    """n >= 2048 takes the cell-grid path (csrc/ball_query_grid.cu); every case must equal the plain scan."""
    from ws3d_b200 import pointnet2_utils
    rng = np.random.default_rng(len(kind) * 977)
    b, n, m, r, k = 2, 4096, 512, 0.5, 32
    xyz = rng.uniform(-20, 20, (b, n, 3)).astype(np.float32)
    if kind == "dense_cluster":      # > 128 hits per ball: hit-buffer overflow -> in-order early-exit scan
        xyz[:, : n // 2] = rng.normal(0, 0.15, (b, n // 2, 3)).astype(np.float32)
    elif kind == "duplicates":
        xyz[:, 1::2] = xyz[:, 0::2]
    elif kind == "flat":             # one cell layer in y
        xyz[..., 1] = 1.5
    elif kind == "line":             # 1-D: per-axis cell cap and enlargement of the cell edge
        xyz[..., 1:] = 0.0
        xyz[..., 0] = rng.uniform(-3000, 3000, (b, n)).astype(np.float32)
    elif kind == "nonfinite":
        xyz[:, 5, 0] = np.nan; xyz[:, 9, 1] = np.inf; xyz[:, 11, 2] = -np.inf
    elif kind == "tiny_radius":
        r = 1e-4
    elif kind == "huge_radius":      # a single cell: everything is a candidate
        r, k = 100.0, 16
    elif kind == "offset":           # large coordinates: rounding of the cell coordinate
        xyz += np.float32(5000.0)
    elif kind == "integer_lattice":  # many points exactly on cell boundaries and at distance exactly r
        xyz = rng.integers(-8, 9, (b, n, 3)).astype(np.float32) * np.float32(0.5)
    pick = rng.integers(0, n, (b, m))
    new_xyz = np.take_along_axis(xyz, pick[..., None], 1).copy()
    new_xyz[:, :16] += np.float32(0.3)           # centres that are not points
    new_xyz[:, 16:24] += np.float32(1000.0)      # far outside the bounding box
    new_xyz[:, 24:28, 0] -= np.float32(0.4)      # just outside / on the box faces
    if kind == "nonfinite":
        new_xyz[:, 30, 0] = np.nan; new_xyz[:, 31, 2] = np.inf
    with np.errstate(invalid="ignore", over="ignore"):
        exp = oracle.ball_query(r, k, xyz, new_xyz)
    x, q = _t(xyz), _t(new_xyz)
    idx = pointnet2_utils.ball_query(r, k, x, q)
    np.testing.assert_array_equal(idx.cpu().numpy(), exp)
    i0, i1 = pointnet2_utils.ball_query_pair((r * 0.5, r), (16, k), x, q)
    np.testing.assert_array_equal(i1.cpu().numpy(), exp)
    with np.errstate(invalid="ignore", over="ignore"):
        np.testing.assert_array_equal(i0.cpu().numpy(), oracle.ball_query(r * 0.5, 16, xyz, new_xyz))
    ref = require_ref("pointnet2_cuda")
    if ref is not None:
        ridx = torch.zeros((b, m, k), dtype=torch.int32, device=dev)
        ref.ball_query_wrapper(b, n, m, r, k, q, x, ridx)
        assert torch.equal(ridx, idx)


def test_ball_query_pair_equals_two_queries():
    from ws3d_b200 import pointnet2_utils
    rng = np.random.default_rng(5)
    xyz = _cloud(rng, 2, 8192, "scene")
    x = _t(xyz)
    _, q = pointnet2_utils.sample_and_gather(x, 512)
    i0, i1 = pointnet2_utils.ball_query_pair((0.5, 1.0), (16, 32), x, q)
    assert torch.equal(i0, pointnet2_utils.ball_query(0.5, 16, x, q))
    assert torch.equal(i1, pointnet2_utils.ball_query(1.0, 32, x, q))


@pytest.mark.parametrize("b,c,n,m,k", [(2, 3, 1024, 256, 16), (1, 96, 4096, 1024, 32), (2, 5, 333, 50, 7), (1, 1, 16384, 4096, 32)])
def test_group_and_gather(b, c, n, m, k):
    from ws3d_b200 import pointnet2_utils
    rng = np.random.default_rng(c + n)
    feat = rng.normal(size=(b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, (b, m, k)).astype(np.int32)
    f, i = _t(feat).requires_grad_(True), _t(idx)
    out = pointnet2_utils.grouping_operation(f, i)
    np.testing.assert_array_equal(out.detach().cpu().numpy(), oracle.grouping_operation(feat, idx))
    g = rng.normal(size=out.shape).astype(np.float32)
    out.backward(_t(g))
    np.testing.assert_allclose(f.grad.cpu().numpy(), oracle.grouping_operation_grad(g, idx, n), rtol=1e-5, atol=1e-5)
    # gather == group with one sample per row
    idx1 = np.ascontiguousarray(idx[:, :, 0])
    f2 = _t(feat).requires_grad_(True)
    out1 = pointnet2_utils.gather_operation(f2, _t(idx1))
    np.testing.assert_array_equal(out1.detach().cpu().numpy(), oracle.gather_operation(feat, idx1))
    g1 = rng.normal(size=out1.shape).astype(np.float32)
    out1.backward(_t(g1))
    np.testing.assert_allclose(f2.grad.cpu().numpy(), oracle.gather_operation_grad(g1, idx1, n), rtol=1e-5, atol=1e-5)
    ref = require_ref("pointnet2_cuda")
    if ref is not None:
        r = torch.empty_like(out)
        ref.group_points_wrapper(b, c, n, m, k, f.detach(), i, r)
        assert torch.equal(r, out.detach())


@pytest.mark.parametrize("b,n,m", [(2, 4096, 1024), (1, 16384, 4096), (2, 256, 64), (1, 1000, 333), (2, 10, 2), (1, 7, 1),
                                   (1, 100, 5000)])
def test_three_nn(b, n, m):
    from ws3d_b200 import native, pointnet2_utils
    rng = np.random.default_rng(n + m)
    known = _cloud(rng, b, m, "grid" if m == 333 else "uniform")
    unknown = _cloud(rng, b, n, "grid" if m == 333 else "uniform")
    d2, idx = oracle.three_nn(unknown, known)
    u, k = _t(unknown), _t(known)
    gd2 = torch.empty((b, n, 3), device=dev)
    gidx = torch.empty((b, n, 3), dtype=torch.int32, device=dev)
    native.three_nn_wrapper(b, n, m, u, k, gd2, gidx)
    np.testing.assert_array_equal(gidx.cpu().numpy(), idx)
    np.testing.assert_array_equal(gd2.cpu().numpy(), d2)
    dist, idx2 = pointnet2_utils.three_nn(u, k)
    np.testing.assert_array_equal(idx2.cpu().numpy(), idx)
    np.testing.assert_allclose(dist.cpu().numpy(), np.sqrt(d2), rtol=1e-6)
    ref = require_ref("pointnet2_cuda")
    if ref is not None:
        rd2, ridx = torch.empty_like(gd2), torch.empty_like(gidx)
        ref.three_nn_wrapper(b, n, m, u, k, rd2, ridx)
        assert torch.equal(ridx, gidx) and torch.equal(rd2, gd2)


NN_GRID_CASES = ["scene_subset", "lattice_ties", "duplicates", "outliers", "nonfinite", "line", "two_known_cells", "offset"]


@pytest.mark.parametrize("kind", NN_GRID_CASES)
def test_three_nn_cell_grid_edge_cases(kind):
    """m >= 512 and n >= 1024 take the ring search over a cell grid (csrc/three_nn_grid.cu)."""
    from ws3d_b200 import native
    rng = np.random.default_rng(len(kind) * 131 + 7)
    b, n, m = 2, 4096, 1024
    unknown = rng.uniform(-20, 20, (b, n, 3)).astype(np.float32)
    known = rng.uniform(-20, 20, (b, m, 3)).astype(np.float32)
    if kind == "scene_subset":        # FP-layer geometry: known points are a subset of the unknown ones
        from ws3d_b200 import synth
        unknown = synth.make_batch(b, n)[..., :3].copy()
        known = unknown[:, rng.permutation(n)[:m]].copy()
    elif kind == "lattice_ties":      # many exactly equal distances: the index order decides
        unknown = rng.integers(-6, 7, (b, n, 3)).astype(np.float32)
        known = rng.integers(-6, 7, (b, m, 3)).astype(np.float32)
    elif kind == "duplicates":
        known[:, 1::2] = known[:, 0::2]
    elif kind == "outliers":          # queries far outside the known points' box; two isolated known points
        unknown[:, :64] += np.float32(500.0)
        known[:, 0] = np.float32(-900.0); known[:, 1] = np.float32(900.0)
    elif kind == "nonfinite":
        known[:, 3, 0] = np.nan; known[:, 4, 1] = np.inf
        unknown[:, 5, 2] = np.nan; unknown[:, 6, 0] = -np.inf
    elif kind == "line":
        known[..., 1:] = 0.0
        known[..., 0] = rng.uniform(-3000, 3000, (b, m)).astype(np.float32)
    elif kind == "two_known_cells":   # all known points in two tight clumps: long ring walks in between
        known[:, : m // 2] = rng.normal(-15, 0.01, (b, m // 2, 3)).astype(np.float32)
        known[:, m // 2:] = rng.normal(15, 0.01, (b, m - m // 2, 3)).astype(np.float32)
    elif kind == "offset":
        unknown += np.float32(7000.0); known += np.float32(7000.0)
    with np.errstate(invalid="ignore", over="ignore"):
        d2, idx = oracle.three_nn(unknown, known)
    u, k = _t(unknown), _t(known)
    gd2 = torch.empty((b, n, 3), device=dev)
    gidx = torch.empty((b, n, 3), dtype=torch.int32, device=dev)
    native.three_nn_wrapper(b, n, m, u, k, gd2, gidx)
    np.testing.assert_array_equal(gidx.cpu().numpy(), idx)
    np.testing.assert_array_equal(gd2.cpu().numpy(), d2)
    ref = require_ref("pointnet2_cuda")
    if ref is not None:
        rd2, ridx = torch.empty_like(gd2), torch.empty_like(gidx)
        ref.three_nn_wrapper(b, n, m, u, k, rd2, ridx)
        assert torch.equal(ridx, gidx) and torch.equal(rd2, gd2)


@pytest.mark.parametrize("b,c,m,n", [(2, 64, 256, 1024), (1, 256, 4096, 16384), (2, 7, 33, 100)])
def test_three_interpolate(b, c, m, n):
    from ws3d_b200 import pointnet2_utils
    rng = np.random.default_rng(c + m)
    feat = rng.normal(size=(b, c, m)).astype(np.float32)
    idx = rng.integers(0, m, (b, n, 3)).astype(np.int32)
    w = rng.uniform(0, 1, (b, n, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    f = _t(feat).requires_grad_(True)
    out = pointnet2_utils.three_interpolate(f, _t(idx), _t(w))
    np.testing.assert_array_equal(out.detach().cpu().numpy(), oracle.three_interpolate(feat, idx, w))  # same FMA order
    g = rng.normal(size=out.shape).astype(np.float32)
    out.backward(_t(g))
    np.testing.assert_allclose(f.grad.cpu().numpy(), oracle.three_interpolate_grad(g, idx, w, m), rtol=1e-4, atol=1e-5)
    ref = require_ref("pointnet2_cuda")
    if ref is not None:
        r = torch.empty_like(out)
        ref.three_interpolate_wrapper(b, c, m, n, f.detach(), _t(idx), _t(w), r)
        assert torch.equal(r, out.detach())


@pytest.mark.parametrize("b,c,m,n,stencil", [(2, 64, 256, 1024, "nn"), (1, 19, 4096, 16384, "nn"), (2, 7, 33, 100, "random"),
                                             (1, 40, 64, 256, "nn"), (1, 5, 2, 50, "random"), (2, 130, 1024, 4096, "nn"),
                                             (1, 3, 30000, 64, "random")])
def test_three_interpolate_grad_gather_over_inverse_stencil(b, c, m, n, stencil):
    """three_interpolate_grad as a gather over the per-call inverse of the stencil (csrc/interp_grad.cu): equal to the oracle's
    scatter (interpolate_gpu.cu:112-137) up to FP32 summation order, ADDS to grad_points like the reference's atomicAdd,
    and -- unlike the atomic formulation -- bit-reproducible.  Shapes: the four FP levels' row lengths (chunked rows, long rows
    that leave one CTA per SM, channel counts that do not divide the chunk), n not a multiple of 4 (no TMA tail), duplicate
    (i, j) pairs and m < 3, and m beyond the builder's limit (the atomic kernel takes it)."""
    from ws3d_b200 import native
    rng = np.random.default_rng(b * 1000 + c + m)
    if stencil == "nn":       # a real stencil: three nearest known points, inverse-distance weights
        known = _cloud(rng, b, m, "scene")
        unknown = _cloud(rng, b, n, "scene")
        d2, idx = oracle.three_nn(unknown, known)
        r = 1.0 / (np.sqrt(d2) + 1e-8)
        w = (r / r.sum(-1, keepdims=True)).astype(np.float32)
    else:
        idx = rng.integers(0, m, (b, n, 3)).astype(np.int32)
        w = rng.uniform(0, 1, (b, n, 3)).astype(np.float32)
    g = rng.normal(size=(b, c, n)).astype(np.float32)
    want = oracle.three_interpolate_grad(g, idx, w, m)
    outs = []
    for _ in range(2):
        gp = torch.ones((b, c, m), dtype=torch.float32, device=dev)
        native.three_interpolate_grad_wrapper(b, c, n, m, _t(g), _t(idx), _t(w), gp)
        outs.append(gp)
    np.testing.assert_allclose(outs[0].cpu().numpy() - 1.0, want, rtol=1e-4, atol=2e-5)
    if m <= 24576:
        assert torch.equal(outs[0], outs[1])          # fixed summation order
    ref = require_ref("pointnet2_cuda")
    rp = torch.zeros((b, c, m), dtype=torch.float32, device=dev)
    ref.three_interpolate_grad_wrapper(b, c, n, m, _t(g), _t(idx), _t(w), rp)
    np.testing.assert_allclose(outs[0].cpu().numpy() - 1.0, rp.cpu().numpy(), rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("c,use_xyz", [(1, True), (96, True), (0, True), (5, False)])
def test_query_and_group_fused_equals_composition(c, use_xyz):
    from ws3d_b200 import pointnet2_utils
    rng = np.random.default_rng(c)
    b, n, m, k, r = 2, 4096, 512, 32, 1.0
    xyz = _cloud(rng, b, n, "scene")
    feat = rng.normal(size=(b, c, n)).astype(np.float32) if c else None
    x = _t(xyz)
    _, q = pointnet2_utils.sample_and_gather(x, m)
    f = _t(feat).requires_grad_(True) if c else None
    grouper = pointnet2_utils.QueryAndGroup(r, k, use_xyz=use_xyz)
    out = grouper(x, q, f)
    exp = oracle.query_and_group(r, k, xyz, q.cpu().numpy(), feat, use_xyz)
    np.testing.assert_array_equal(out.detach().cpu().numpy(), exp)
    if c:
        g = rng.normal(size=out.shape).astype(np.float32)
        out.backward(_t(g))
        idx = oracle.ball_query(r, k, xyz, q.cpu().numpy())
        off = 3 if use_xyz else 0
        np.testing.assert_allclose(f.grad.cpu().numpy(), oracle.grouping_operation_grad(g[:, off:], idx, n),
                                   rtol=1e-4, atol=1e-5)


def test_full_size_properties():
    """BASELINE config-2 sizes (B=16): checked through size-independent properties."""
    from ws3d_b200 import pointnet2_utils, synth
    pts = torch.from_numpy(synth.make_batch(16)).to(dev)
    xyz = pts[..., :3].contiguous()
    idx, new_xyz = pointnet2_utils.sample_and_gather(xyz, 4096)
    i64 = idx.long()
    assert int(i64.min()) >= 0 and int(i64.max()) < 16384
    assert all(torch.unique(i64[b]).numel() == 4096 for b in range(16))  # distinct points: no repeats
    assert torch.equal(new_xyz, torch.gather(xyz, 1, i64[..., None].expand(-1, -1, 3)))
    assert torch.equal(idx, pointnet2_utils.furthest_point_sample(xyz, 4096))  # idempotent / deterministic
    # FPS prefix property: sampling fewer points gives a prefix
    assert torch.equal(pointnet2_utils.furthest_point_sample(xyz, 512), idx[:, :512].contiguous())
    bq = pointnet2_utils.ball_query(0.5, 32, xyz, new_xyz).long()
    nb = torch.gather(xyz, 1, bq.reshape(16, -1, 1).expand(-1, -1, 3)).reshape(16, 4096, 32, 3)
    d2 = ((nb - new_xyz[:, :, None]) ** 2).sum(-1)
    assert float(d2.max()) < 0.5 * 0.5 * (1 + 1e-5)         # every returned neighbour is inside the ball
    filled = bq[:, :, 1:] >= bq[:, :, :-1]
    first_rep = bq[:, :, 1:] == bq[:, :, :1]
    assert bool((filled | first_rep).all())                  # ascending until the first-hit padding starts
