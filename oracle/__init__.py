"""TEST INFRASTRUCTURE ONLY: numpy front-end of the CPU oracle (oracle/ws3d_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product package (ws3d_b200/) never does.

Each function mirrors one reference entry point (file:line cited in ws3d_oracle.c),
allocates its outputs exactly as the reference's Python wrappers do (zero-filled
idx for ball_query, 1e10-filled temp for FPS, ...) and returns numpy arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libws3d_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile libws3d_oracle.so in place (gcc, seconds)."""
    src = os.path.join(_HERE, "ws3d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libws3d_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_nms.restype = ctypes.c_int
        _lib.oracle_nms_normal.restype = ctypes.c_int
        _lib.oracle_opt_n_threads.restype = ctypes.c_int
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def opt_n_threads(n: int) -> int:
    return lib().oracle_opt_n_threads(int(n))


def furthest_point_sample(xyz, npoint, return_temp=False):
    xyz = _f(xyz)
    b, n, _ = xyz.shape
    temp = np.full((b, n), 1e10, dtype=np.float32)
    idx = np.zeros((b, npoint), dtype=np.int32)
    lib().oracle_furthest_point_sampling(b, n, int(npoint), _p(xyz), _p(temp), _p(idx))
    return (idx, temp) if return_temp else idx


def gather_operation(features, idx):
    features, idx = _f(features), _i(idx)
    b, c, n = features.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), dtype=np.float32)
    lib().oracle_gather_points(b, c, n, m, _p(features), _p(idx), _p(out))
    return out


def gather_operation_grad(grad_out, idx, n):
    grad_out, idx = _f(grad_out), _i(idx)
    b, c, m = grad_out.shape
    g = np.zeros((b, c, n), dtype=np.float32)
    lib().oracle_gather_points_grad(b, c, int(n), m, _p(grad_out), _p(idx), _p(g))
    return g


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((b, m, nsample), dtype=np.int32)
    lib().oracle_ball_query(b, n, m, ctypes.c_float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx))
    return idx


def grouping_operation(features, idx):
    features, idx = _f(features), _i(idx)
    b, c, n = features.shape
    _, m, k = idx.shape
    out = np.empty((b, c, m, k), dtype=np.float32)
    lib().oracle_group_points(b, c, n, m, k, _p(features), _p(idx), _p(out))
    return out


def grouping_operation_grad(grad_out, idx, n):
    grad_out, idx = _f(grad_out), _i(idx)
    b, c, m, k = grad_out.shape
    g = np.zeros((b, c, n), dtype=np.float32)
    lib().oracle_group_points_grad(b, c, int(n), m, k, _p(grad_out), _p(idx), _p(g))
    return g


def three_nn(unknown, known):
    """Returns (dist2, idx): SQUARED distances like the native entry point
    (the reference's Python wrapper applies sqrt afterwards)."""
    unknown, known = _f(unknown), _f(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = np.empty((b, n, 3), dtype=np.float32)
    idx = np.empty((b, n, 3), dtype=np.int32)
    lib().oracle_three_nn(b, n, m, _p(unknown), _p(known), _p(dist2), _p(idx))
    return dist2, idx


def three_interpolate(features, idx, weight):
    features, idx, weight = _f(features), _i(idx), _f(weight)
    b, c, m = features.shape
    n = idx.shape[1]
    out = np.empty((b, c, n), dtype=np.float32)
    lib().oracle_three_interpolate(b, c, m, n, _p(features), _p(idx), _p(weight), _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f(grad_out), _i(idx), _f(weight)
    b, c, n = grad_out.shape
    g = np.zeros((b, c, m), dtype=np.float32)
    lib().oracle_three_interpolate_grad(b, c, n, int(m), _p(grad_out), _p(idx), _p(weight), _p(g))
    return g


def query_and_group(radius, nsample, xyz, new_xyz, features=None, use_xyz=True):
    """pointnet2_utils.py:241-264 (QueryAndGroup.forward) composed from the oracle ops."""
    idx = ball_query(radius, nsample, xyz, new_xyz)
    xyz_t = np.ascontiguousarray(np.transpose(_f(xyz), (0, 2, 1)))
    grouped_xyz = grouping_operation(xyz_t, idx)
    grouped_xyz = grouped_xyz - np.transpose(_f(new_xyz), (0, 2, 1))[..., None]
    if features is None:
        return grouped_xyz
    g = grouping_operation(features, idx)
    return np.concatenate([grouped_xyz, g], axis=1) if use_xyz else g


def boxes_overlap_bev(boxes_a, boxes_b):
    a, b = _f(boxes_a), _f(boxes_b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().oracle_boxes_overlap_bev(a.shape[0], _p(a), b.shape[0], _p(b), _p(out))
    return out


def boxes_iou_bev(boxes_a, boxes_b):
    a, b = _f(boxes_a), _f(boxes_b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().oracle_boxes_iou_bev(a.shape[0], _p(a), b.shape[0], _p(b), _p(out))
    return out


def nms(boxes_sorted, thresh):
    """boxes already sorted by descending score; returns int64 keep indices."""
    bx = _f(boxes_sorted)
    keep = np.zeros(bx.shape[0], dtype=np.int64)
    n = lib().oracle_nms(_p(bx), bx.shape[0], ctypes.c_float(thresh), _p(keep))
    return keep[:n]


def nms_normal(boxes_sorted, thresh):
    bx = _f(boxes_sorted)
    keep = np.zeros(bx.shape[0], dtype=np.int64)
    n = lib().oracle_nms_normal(_p(bx), bx.shape[0], ctypes.c_float(thresh), _p(keep))
    return keep[:n]


def roipool3d(pts, pts_feature, boxes3d, sampled_pt_num=512):
    """Native-level semantics (no box enlargement): returns (pooled, empty_flag)."""
    pts, pts_feature, boxes3d = _f(pts), _f(pts_feature), _f(boxes3d)
    b, n, _ = pts.shape
    m = boxes3d.shape[1]
    c = pts_feature.shape[2]
    pooled = np.zeros((b, m, sampled_pt_num, 3 + c), dtype=np.float32)
    flag = np.zeros((b, m), dtype=np.int32)
    lib().oracle_roipool3d(b, n, m, c, int(sampled_pt_num), _p(pts), _p(boxes3d), _p(pts_feature),
                           _p(pooled), _p(flag))
    return pooled, flag


def pts_in_boxes3d_cpu(pts, boxes3d):
    pts, boxes3d = _f(pts), _f(boxes3d)
    flag = np.zeros((boxes3d.shape[0], pts.shape[0]), dtype=np.int64)
    lib().oracle_pts_in_boxes3d_cpu(_p(flag), _p(pts), _p(boxes3d), boxes3d.shape[0], pts.shape[0])
    return flag


def roipool3d_cpu(pts, boxes3d, pts_feature, sampled_pt_num):
    pts, boxes3d, pts_feature = _f(pts), _f(boxes3d), _f(pts_feature)
    m, c = boxes3d.shape[0], pts_feature.shape[1]
    pooled_pts = np.zeros((m, sampled_pt_num, 3), dtype=np.float32)
    pooled_feat = np.zeros((m, sampled_pt_num, c), dtype=np.float32)
    flag = np.zeros(m, dtype=np.int64)
    lib().oracle_roipool3d_cpu(_p(pts), _p(boxes3d), _p(pts_feature), _p(pooled_pts), _p(pooled_feat),
                               _p(flag), m, pts.shape[0], c, int(sampled_pt_num))
    return pooled_pts, pooled_feat, flag


# ---- SURVEY.md section 8 "next" rows f2 / f3 ---------------------------------------------------
def boxes_iou3d_aligned(boxes_a, boxes_b):
    """Diagonal of boxes_iou3d_gpu (iou3d_utils.py:21-56): (n,7),(n,7) -> iou2d (n), iou3d (n)."""
    a, b = _f(boxes_a), _f(boxes_b)
    assert a.shape == b.shape and a.shape[1] == 7
    i2, i3 = np.zeros(a.shape[0], dtype=np.float32), np.zeros(a.shape[0], dtype=np.float32)
    lib().oracle_boxes_iou3d_aligned(a.shape[0], _p(a), _p(b), _p(i2), _p(i3))
    return i2, i3


def radius_nms(centers_sorted, radius):
    """tools/eval_auto.py:263-279: centres (n,2) sorted by descending score -> int64 keep indices."""
    c = _f(centers_sorted)
    keep = np.zeros(c.shape[0], dtype=np.int64)
    lib().oracle_radius_nms.restype = ctypes.c_int
    n = lib().oracle_radius_nms(_p(c), c.shape[0], ctypes.c_float(radius), _p(keep))
    return keep[:n]


def cylinder_query(pts, centers, radius, cap):
    """tools/eval_auto.py:289-291,:327-343: pts (n,3), centres (m,2) -> idx (m,cap) (-1 padded), cnt (m), any (n)."""
    p, c = _f(pts), _f(centers)
    idx = np.full((c.shape[0], cap), -1, dtype=np.int32)
    cnt = np.zeros(c.shape[0], dtype=np.int32)
    any_ = np.zeros(p.shape[0], dtype=np.uint8)
    lib().oracle_cylinder_query(p.shape[0], c.shape[0], int(cap), ctypes.c_float(radius), _p(p), _p(c), _p(idx), _p(cnt),
                                _p(any_))
    return idx, cnt, any_


def gaussian_rpn_labels(pts_rect, gt_boxes3d, gauss_height=0.707, gauss_status=0.7, gauss_cov=1.5, fg_radius=4.0):
    """kitti_rcnn_dataset.py:529-573 for one scene: pts (n,3), boxes (g,7) -> cls (n) float64, reg (n,3) float32."""
    p, b = _f(pts_rect), _f(gt_boxes3d).reshape(-1, 7)
    cls = np.zeros(p.shape[0], dtype=np.float64)
    reg = np.zeros((p.shape[0], 3), dtype=np.float32)
    lib().oracle_gaussian_rpn_labels(p.shape[0], b.shape[0], _p(p), _p(b), ctypes.c_float(gauss_height), ctypes.c_float(gauss_status),
                                     ctypes.c_double(gauss_cov), ctypes.c_float(fg_radius), _p(cls), _p(reg))
    return cls, reg


# ---- SURVEY.md section 8 rows f2 (corner-loss box math) / f4 (loader subsampling) ----------------
def boxes3d_to_corners3d(boxes3d, flip=False):
    """kitti_utils.py:104-131: (n,7) -> (n,8,3)."""
    b = _f(boxes3d).reshape(-1, 7)
    out = np.zeros((b.shape[0], 8, 3), dtype=np.float32)
    lib().oracle_boxes3d_to_corners3d(b.shape[0], _p(b), int(bool(flip)), _p(out))
    return out


def corner_distance(pred_boxes3d, gt_boxes3d):
    """train_functions.py:266-271: aligned (n,7) pairs -> (n,8) min(|P - G|, |P - G_flipped|)."""
    p, g = _f(pred_boxes3d).reshape(-1, 7), _f(gt_boxes3d).reshape(-1, 7)
    out = np.zeros((p.shape[0], 8), dtype=np.float32)
    lib().oracle_corner_distance(p.shape[0], _p(p), _p(g), _p(out))
    return out


def subsample_points(pts, depth, npoints, rng):
    """kitti_rcnn_dataset.py:424-444, restated line by line with `rng` (a numpy RandomState) in place of the global
    np.random: pts (n, 3 + C) rows [xyz, intensity...] -> (pts_input (npoints, 3 + C) with the last channel shifted by
    -0.5 when C > 0, choice (npoints))."""
    pts = _f(pts)
    n = pts.shape[0]
    if npoints < n:
        pts_near_flag = np.asarray(depth, dtype=np.float32) < 40.0
        far_idxs_choice = np.where(pts_near_flag == 0)[0]
        near_idxs = np.where(pts_near_flag == 1)[0]
        near_idxs_choice = rng.choice(near_idxs, npoints - len(far_idxs_choice), replace=False)
        choice = np.concatenate((near_idxs_choice, far_idxs_choice), axis=0) if len(far_idxs_choice) > 0 else near_idxs_choice
        rng.shuffle(choice)
    else:
        choice = np.arange(0, n, dtype=np.int32)
        extra_choice = np.arange(0, n, dtype=np.int32)
        while npoints > len(choice):
            choice = np.concatenate((choice, extra_choice), axis=0)
        choice = rng.choice(choice, npoints, replace=False)
        rng.shuffle(choice)
    out = pts[choice, :].copy()
    if pts.shape[1] > 3:
        out[:, -1] = out[:, -1] - np.float32(0.5)
    return out, choice.astype(np.int32)
