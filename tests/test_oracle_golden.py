"""CPU: the oracle is pinned to outputs of the reference's OWN kernels (tests/golden/*.npz, produced
on a B200 by tools/make_golden.py from oracle/_ref, i.e. the unmodified reference sources).
Index-producing ops and copies: bit-exact.  iou3d values: 1e-5 (host libm vs libdevice trig)."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run tools/make_golden.py on the GPU box and commit the fixtures")
    return np.load(path)


def test_fps_golden():
    g = _load("fps.npz")
    for name in ("scene", "uniform", "grid", "small", "big"):
        xyz, idx, temp = g[name + "_xyz"], g[name + "_idx"], g[name + "_temp"]
        oi, ot = oracle.furthest_point_sample(xyz, idx.shape[1], return_temp=True)
        np.testing.assert_array_equal(oi, idx, err_msg=name)
        np.testing.assert_array_equal(ot, temp, err_msg=name)


def test_pointnet2_golden():
    g = _load("pointnet2.npz")
    xyz, new_xyz = g["xyz"], g["new_xyz"]
    for r, k in ((0.5, 16), (1.0, 32), (4.0, 8)):
        np.testing.assert_array_equal(oracle.ball_query(r, k, xyz, new_xyz), g[f"bq_{r}_{k}"])
    np.testing.assert_array_equal(oracle.grouping_operation(g["feat"], g["bq_1.0_32"]), g["grouped"])
    np.testing.assert_array_equal(oracle.gather_operation(g["feat"], g["fps_idx"]), g["gathered"])
    d2, idx = oracle.three_nn(xyz, g["known"])
    np.testing.assert_array_equal(idx, g["nn_idx"])
    np.testing.assert_array_equal(d2, g["nn_dist2"])
    np.testing.assert_array_equal(oracle.three_interpolate(g["kfeat"], g["nn_idx"], g["weight"]), g["interp"])


def test_iou3d_golden():
    g = _load("iou3d.npz")
    np.testing.assert_allclose(oracle.boxes_overlap_bev(g["a"], g["b"]), g["overlap"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle.boxes_iou_bev(g["a"], g["b"]), g["iou"], rtol=1e-5, atol=1e-5)
    for th in (0.1, 0.5, 0.85):
        np.testing.assert_array_equal(oracle.nms_normal(g["nms_boxes"], th), g[f"nmsn_{th}"])
        ref_keep, ours = g[f"nms_{th}"], oracle.nms(g["nms_boxes"], th)
        # rotated NMS goes through cosf/sinf/atan2f: identical unless a pair sits within 1 ulp of thresh
        assert len(set(ref_keep) ^ set(ours)) <= 2, (th, len(ref_keep), len(ours))


def test_roipool3d_golden():
    g = _load("roipool3d.npz")
    pooled, flag = oracle.roipool3d(g["xyz"], g["feat"], g["boxes"], g["pooled"].shape[2])
    np.testing.assert_array_equal(flag, g["flag"])
    np.testing.assert_array_equal(pooled, g["pooled"])


def test_next_rows_f2_f3_match_reference_python():
    """the oracle restatements against vectors produced by executing the reference's
    own Python (tools/make_golden_next.py: boxes_iou3d_gpu diagonal, distance_2-driven radius NMS + cylinder crop)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "next_rows.npz"))
    i2, i3 = oracle.boxes_iou3d_aligned(g["f2_boxes_a"], g["f2_boxes_b"])
    np.testing.assert_array_equal(i2, g["f2_iou2d_diag"])
    np.testing.assert_array_equal(i3, g["f2_iou3d_diag"])
    assert (g["f2_iou3d_diag"][:40] > 0.999).all() and (g["f2_iou3d_diag"][40:100] == 0).all()
    centres_sorted = g["f3_centres"][g["f3_sort"]]
    keep = oracle.radius_nms(centres_sorted, 0.3)
    np.testing.assert_array_equal(keep, g["f3_keep_id"])
    assert 1 < len(keep) < len(centres_sorted)
    idx, cnt, any_ = oracle.cylinder_query(g["f3_points"], centres_sorted[keep], 4.0, g["f3_idx"].shape[1])
    np.testing.assert_array_equal(cnt, g["f3_cnt"])
    np.testing.assert_array_equal(idx, g["f3_idx"])
    np.testing.assert_array_equal(any_, g["f3_any"])


def test_next_row_f4_gaussian_labels_match_reference_python():
    """oracle_gaussian_rpn_labels against the output of the reference's own
    generate_gaussian_training_labels (executed from its source text by tools/make_golden_next.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "next_rows.npz"))
    cls, reg = oracle.gaussian_rpn_labels(g["f4_points"], g["f4_boxes"])
    np.testing.assert_array_equal(reg, g["f4_reg"])                        # float32 arithmetic: bit-exact
    np.testing.assert_allclose(cls, g["f4_cls"], rtol=0, atol=1e-12)       # scipy's float64 Gaussian vs exp(-d^2 / 2 cov)
    assert 0 < int((reg[:, 0] != 0).sum()) < reg.shape[0] and float(cls.max()) == 1.0
    cls0, reg0 = oracle.gaussian_rpn_labels(g["f4_points"][:100], g["f4_boxes"][:0])
    np.testing.assert_array_equal(cls0, g["f4_cls_empty"])
    np.testing.assert_array_equal(reg0, g["f4_reg_empty"])
