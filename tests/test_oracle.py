"""CPU: known-answer and property tests of the oracle itself (oracle/ws3d_oracle.c)."""
import numpy as np
import pytest

import oracle


def test_opt_n_threads_matches_reference_formula():
    # cuda_utils.h:10-14: largest power of two <= n, capped at 1024 (float log quirks included)
    for n, exp in [(1, 1), (2, 2), (3, 2), (100, 64), (255, 128), (256, 256), (512, 512), (1000, 512), (1024, 1024),
                   (4096, 1024), (16384, 1024)]:
        assert oracle.opt_n_threads(n) == exp


def test_fps_basic_and_tie_break_rule():
    # 4 points on a line: 0, 1, 10, 4 -> picks 0, then the farthest (10), then 4 (min-dist 4 vs 1)
    xyz = np.array([[[0, 0, 0], [1, 0, 0], [10, 0, 0], [4, 0, 0]]], np.float32)
    np.testing.assert_array_equal(oracle.furthest_point_sample(xyz, 3)[0], [0, 2, 3])
    # exact tie: points 1 and 2 are both at distance 5 from point 0.  Block size for n=4 is 4; the
    # shared-memory tree keeps the LOWER slot of each pair (tid, tid+2) then (0,1): the winner is
    # the candidate whose bit-reversed thread id is smallest: tid 2 (bits 10 -> 01) beats tid 1 (01 -> 10).
    xyz = np.array([[[0, 0, 0], [5, 0, 0], [0, 5, 0], [1, 1, 0]]], np.float32)
    assert oracle.furthest_point_sample(xyz, 2)[0, 1] == 2
    # same geometry with the tied points at indices 1 and 3: tid 1 (rev 2) beats tid 3 (rev 3)
    xyz = np.array([[[0, 0, 0], [5, 0, 0], [1, 1, 0], [0, 5, 0]]], np.float32)
    assert oracle.furthest_point_sample(xyz, 2)[0, 1] == 1


def test_fps_more_samples_than_points_and_temp():
    xyz = np.random.default_rng(0).uniform(-1, 1, (2, 16, 3)).astype(np.float32)
    idx, temp = oracle.furthest_point_sample(xyz, 24, return_temp=True)
    assert sorted(idx[0, :16].tolist()) == list(range(16))     # all points once ...
    assert (idx[:, 16:] == 0).all() and (temp == 0).all()      # ... then index 0 forever (all distances 0)


def test_ball_query_semantics():
    xyz = np.array([[[0, 0, 0], [0.5, 0, 0], [3, 0, 0], [0.1, 0, 0], [0.2, 0, 0]]], np.float32)
    new_xyz = np.array([[[0, 0, 0], [100, 0, 0]]], np.float32)
    idx = oracle.ball_query(1.0, 3, xyz, new_xyz)
    np.testing.assert_array_equal(idx[0, 0], [0, 1, 3])          # first 3 hits in index order
    np.testing.assert_array_equal(idx[0, 1], [0, 0, 0])          # no hit: caller's zero fill
    idx = oracle.ball_query(0.15, 4, xyz, new_xyz)
    np.testing.assert_array_equal(idx[0, 0], [0, 3, 0, 0])       # fewer than K: first hit pads
    # strict '<' on the f32 product r*r
    p = np.array([[[0, 0, 0], [1, 0, 0]]], np.float32)
    np.testing.assert_array_equal(oracle.ball_query(1.0, 2, p, p[:, :1])[0, 0], [0, 0])


def test_three_nn_order_ties_and_short_inputs():
    known = np.array([[[1, 0, 0], [2, 0, 0], [1, 0, 0], [0.5, 0, 0]]], np.float32)
    unk = np.zeros((1, 1, 3), np.float32)
    d2, idx = oracle.three_nn(unk, known)
    np.testing.assert_array_equal(idx[0, 0], [3, 0, 2])          # ascending distance, lower index first on ties
    np.testing.assert_allclose(d2[0, 0], [0.25, 1, 1])
    d2, idx = oracle.three_nn(unk, known[:, :2])
    assert np.isinf(d2[0, 0, 2]) and idx[0, 0, 2] == 0           # (float)1e40 == inf, index 0


def test_group_gather_interpolate_and_grads_are_adjoint():
    rng = np.random.default_rng(1)
    b, c, n, m, k = 2, 3, 50, 7, 5
    feat = rng.normal(size=(b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, (b, m, k)).astype(np.int32)
    g = rng.normal(size=(b, c, m, k)).astype(np.float32)
    out = oracle.grouping_operation(feat, idx)
    assert out[1, 2, 3, 4] == feat[1, 2, idx[1, 3, 4]]
    lhs = float((out.astype(np.float64) * g).sum())
    rhs = float((feat.astype(np.float64) * oracle.grouping_operation_grad(g, idx, n)).sum())
    assert abs(lhs - rhs) < 1e-3 * max(1, abs(lhs))              # <G(f), g> == <f, G^T(g)>
    w = rng.uniform(0, 1, (b, m, 3)).astype(np.float32)
    i3 = rng.integers(0, n, (b, m, 3)).astype(np.int32)
    gi = rng.normal(size=(b, c, m)).astype(np.float32)
    o = oracle.three_interpolate(feat, i3, w)
    lhs = float((o.astype(np.float64) * gi).sum())
    rhs = float((feat.astype(np.float64) * oracle.three_interpolate_grad(gi, i3, w, n)).sum())
    assert abs(lhs - rhs) < 1e-3 * max(1, abs(lhs))


def test_box_overlap_known_answers():
    sq = np.array([[0, 0, 2, 2, 0.0]], np.float32)
    np.testing.assert_allclose(oracle.boxes_overlap_bev(sq, sq), [[4.0]], rtol=1e-6)
    np.testing.assert_allclose(oracle.boxes_iou_bev(sq, sq), [[1.0]], rtol=1e-5)
    half = np.array([[1, 0, 3, 2, 0.0]], np.float32)                # shifted by half a side: overlap 2, IoU 2/6
    np.testing.assert_allclose(oracle.boxes_overlap_bev(sq, half), [[2.0]], rtol=1e-5)
    np.testing.assert_allclose(oracle.boxes_iou_bev(sq, half), [[1 / 3]], rtol=1e-5)
    rot = np.array([[0, 0, 2, 2, np.pi / 4]], np.float32)           # same square turned 45 deg: octagon
    np.testing.assert_allclose(oracle.boxes_overlap_bev(sq, rot), [[8 * (np.sqrt(2) - 1)]], rtol=1e-5)
    far = np.array([[10, 10, 12, 12, 0.3]], np.float32)
    assert oracle.boxes_overlap_bev(sq, far)[0, 0] == 0.0
    inner = np.array([[0.5, 0.5, 1.5, 1.5, 0.7]], np.float32)       # fully contained, rotated
    np.testing.assert_allclose(oracle.boxes_overlap_bev(sq, inner), [[1.0]], rtol=1e-5)


def test_nms_greedy_semantics():
    boxes = np.array([[0, 0, 2, 2, 0], [0.1, 0, 2.1, 2, 0], [5, 5, 7, 7, 0.2], [0, 0.2, 2, 2.2, 0], [5.05, 5, 7.05, 7, 0.2]], np.float32)
    np.testing.assert_array_equal(oracle.nms(boxes, 0.5), [0, 2])
    np.testing.assert_array_equal(oracle.nms(boxes, 0.99), [0, 1, 2, 3, 4])
    np.testing.assert_array_equal(oracle.nms_normal(boxes, 0.5), [0, 2])
    # suppression is not transitive through suppressed boxes: b suppressed by a cannot suppress c
    chain = np.array([[0, 0, 4, 1, 0], [1.5, 0, 5.5, 1, 0], [3, 0, 7, 1, 0]], np.float32)
    np.testing.assert_array_equal(oracle.nms_normal(chain, 0.3), [0, 2])
    rng = np.random.default_rng(2)
    n = 130                                                          # not a multiple of 64
    c = rng.uniform(0, 10, (n, 2))
    b = np.concatenate([c - 1, c + 1, rng.uniform(-3, 3, (n, 1))], 1).astype(np.float32)
    keep = oracle.nms(b, 0.3)
    iou = oracle.boxes_iou_bev(b[keep], b[keep])
    np.fill_diagonal(iou, 0)
    assert iou.max() <= 0.3 + 1e-6 and len(keep) < n


def test_roipool_truncation_wraparound_and_empty():
    xyz = np.zeros((1, 10, 3), np.float32)
    xyz[0, :, 0] = np.arange(10) * 0.1           # 10 points along x inside a 2 m box
    xyz[0, 7:, 0] += 50                          # last three far away
    feat = np.arange(10, dtype=np.float32).reshape(1, 10, 1)
    box = np.array([[[0.3, 1.0, 0, 2, 2, 2, 0.0], [100, 1, 0, 2, 2, 2, 0]]], np.float32)  # y is the BOTTOM centre
    xyz[0, :, 1] = 0.0
    pooled, flag = oracle.roipool3d(xyz, feat, box, 4)
    np.testing.assert_array_equal(flag, [[0, 1]])
    np.testing.assert_array_equal(pooled[0, 0, :, 3], [0, 1, 2, 3])              # first S in index order
    assert (pooled[0, 1] == 0).all()                                            # empty box untouched
    pooled, flag = oracle.roipool3d(xyz, feat, box, 10)
    np.testing.assert_array_equal(pooled[0, 0, :, 3], [0, 1, 2, 3, 4, 5, 6, 0, 1, 2])  # wrap-around k % cnt


def test_next_rows_oracle_properties():
    """Size-independent properties of the f3 / f4 restatements (no fixtures involved)."""
    rng = np.random.default_rng(5)
    cen = rng.uniform(0, 6, (400, 2)).astype(np.float32)
    keep = oracle.radius_nms(cen, 0.5)
    assert keep[0] == 0 and np.all(np.diff(keep) > 0)
    kept = cen[keep]
    d = np.sqrt(((kept[:, None] - kept[None]) ** 2).sum(-1))
    np.fill_diagonal(d, 9.0)
    assert d.min() > 0.5                                                  # kept centres are pairwise farther than the radius
    dropped = np.setdiff1d(np.arange(400), keep)
    for i in dropped[:50]:                                                # each dropped one has an EARLIER kept centre within it
        earlier = keep[keep < i]
        assert np.sqrt(((cen[earlier] - cen[i]) ** 2).sum(-1)).min() <= 0.5 + 1e-6
    # cylinder membership: counts, order, any-flag
    pts = rng.uniform(-10, 10, (3000, 3)).astype(np.float32)
    idx, cnt, any_ = oracle.cylinder_query(pts, kept[:20] - 3.0, 4.0, 64)
    for c in range(20):
        inside = np.nonzero(np.sqrt(((pts[:, [0, 2]] - (kept[c] - 3.0)) ** 2).sum(-1)) < 4.0)[0]
        assert abs(cnt[c] - len(inside)) <= 1                             # float64 here vs float32 there: boundary ties only
        got = idx[c, :min(cnt[c], 64)]
        assert np.all(np.diff(got) > 0)
    assert any_.sum() > 0
    # Gaussian labels: cls in (0, 1], 1 inside the 0.7 m core, regression target points at the nearest box within 4 m
    boxes = np.array([[0, 1.6, 10, 1.5, 1.6, 3.9, 0.3], [8, 1.6, 30, 1.5, 1.6, 3.9, -1.0]], np.float32)
    p = np.array([[0.1, 0.0, 10.1], [7.0, 0.0, 29.0], [40, 0, 60], [0, 0, 13.5]], np.float32)
    cls, reg = oracle.gaussian_rpn_labels(p, boxes)
    assert cls[0] == 1.0 and 0 < cls[1] < 1 and cls[2] < 1e-100          # 40 m from the nearest box: exp(-d^2 / 3)
    np.testing.assert_allclose(reg[0], [-0.1, 0, -0.1], atol=1e-6)
    np.testing.assert_allclose(reg[1], [1.0, 0, 1.0], atol=1e-6)
    assert np.all(reg[2] == 0) and reg[3, 2] == np.float32(10) - np.float32(13.5)


def _fps_thread_level(xyz, m):
    """A thread-by-thread emulation of furthest_point_sampling_kernel<block_size> (sampling_gpu.cu:93-209) for one cloud:
    strided per-thread scan with strict '>', shared arrays dists / dists_i, the unrolled tree of __update calls
    (:86-91: the lower slot keeps its index unless the upper value is strictly larger), old = dists_i[0].
    Integer-lattice inputs only: every distance is exact in float32, so the FMA shape does not matter here."""
    n = xyz.shape[0]
    bs = oracle.opt_n_threads(n)
    temp = np.full(n, 1e10, np.float32)
    idx = np.zeros(m, np.int32)
    old = 0
    for j in range(1, m):
        dists = np.full(bs, -1.0, np.float32)
        dists_i = np.zeros(bs, np.int64)
        d = ((xyz - xyz[old]) ** 2).sum(1).astype(np.float32)
        temp = np.minimum(d, temp)
        for tid in range(bs):
            best, besti = np.float32(-1), 0
            for k in range(tid, n, bs):
                if temp[k] > best:
                    best, besti = temp[k], k
            dists[tid], dists_i[tid] = best, besti
        s = bs // 2
        while s >= 1:
            for tid in range(s):
                v1, v2 = dists[tid], dists[tid + s]
                if v2 > v1:
                    dists_i[tid] = dists_i[tid + s]
                dists[tid] = max(v1, v2)
            s //= 2
        old = int(dists_i[0])
        idx[j] = old
    return idx


@pytest.mark.parametrize("n,m,span", [(8, 8, 2), (37, 20, 3), (64, 40, 3), (100, 100, 4), (300, 120, 5)])
def test_fps_oracle_equals_thread_level_emulation_on_lattices(n, m, span):
    """Lattice clouds are full of exact ties and duplicates: the oracle's closed-form tie rule (largest value, then the
    smallest bit-reversed k mod BS, then the smallest k) must pick what the reference kernel's shared-memory tree picks."""
    rng = np.random.default_rng(n * 31 + m)
    xyz = rng.integers(-span, span + 1, (n, 3)).astype(np.float32)
    want = _fps_thread_level(xyz, m)
    got = oracle.furthest_point_sample(xyz[None], m)[0]
    np.testing.assert_array_equal(got, want)
