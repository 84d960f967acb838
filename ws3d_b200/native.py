"""Tensor-level entry points with the reference's pybind signatures.

One function per entry of the three reference extension modules
(pointnet2_lib/pointnet2/src/pointnet2_api.cpp:11-23, lib/utils/iou3d/src/iou3d.cpp:175-178,
lib/utils/roipool3d/src/roipool3d.cpp:199-202): same names, same positional arguments, the
caller allocates every output.  They forward raw pointers to libws3d_ops.so on torch's current
stream.  ws3d_b200/dropin/{pointnet2_cuda,iou3d_cuda,roipool3d_cuda}.py re-export them under
the reference module names so the reference's Python wrappers run on top unmodified.
"""
import torch

from . import _C
from ._C import F32, F64, I32, I64, U8, check, device_of, lib, ptr, require, stream


# ---- pointnet2_cuda ---------------------------------------------------------------------------
def furthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    require("furthest_point_sampling", (points_tensor, F32, b * n * 3), (temp_tensor, F32, b * n), (idx_tensor, I32, b * m))
    with device_of(points_tensor):
        check(lib().ws3d_furthest_point_sampling(b, n, m, ptr(points_tensor), ptr(temp_tensor), ptr(idx_tensor),
                                                 stream()), "furthest_point_sampling")
    return 1


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    require("gather_points", (points_tensor, F32, b * c * n), (idx_tensor, I32, b * npoints), (out_tensor, F32, b * c * npoints))
    with device_of(points_tensor):
        check(lib().ws3d_gather_points(b, c, n, npoints, ptr(points_tensor), ptr(idx_tensor), ptr(out_tensor),
                                       stream()), "gather_points")
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    require("gather_points_grad", (grad_out_tensor, F32, b * c * npoints), (idx_tensor, I32, b * npoints),
            (grad_points_tensor, F32, b * c * n))
    with device_of(grad_out_tensor):
        check(lib().ws3d_gather_points_grad(b, c, n, npoints, ptr(grad_out_tensor), ptr(idx_tensor),
                                            ptr(grad_points_tensor), stream()), "gather_points_grad")
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    require("ball_query", (new_xyz_tensor, F32, b * m * 3), (xyz_tensor, F32, b * n * 3), (idx_tensor, I32, b * m * nsample))
    with device_of(xyz_tensor):
        check(lib().ws3d_ball_query(b, n, m, float(radius), nsample, ptr(new_xyz_tensor), ptr(xyz_tensor),
                                    ptr(idx_tensor), stream()), "ball_query")
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    require("group_points", (points_tensor, F32, b * c * n), (idx_tensor, I32, b * npoints * nsample),
            (out_tensor, F32, b * c * npoints * nsample))
    with device_of(points_tensor):
        check(lib().ws3d_group_points(b, c, n, npoints, nsample, ptr(points_tensor), ptr(idx_tensor),
                                      ptr(out_tensor), stream()), "group_points")
    return 1


def group_concat_grad(b, n, m, c, nsample, use_xyz, grad_out, idx, grad_features):
    """Gradient of group_concat with respect to the features, reading grad_out (B, 3*use_xyz + C, M, nsample) in place."""
    lead = 3 if use_xyz else 0
    require("group_concat_grad", (grad_out, F32, b * (lead + c) * m * nsample), (idx, I32, b * m * nsample), (grad_features, F32, b * c * n))
    with device_of(grad_out):
        check(lib().ws3d_group_concat_grad(b, n, m, c, nsample, int(bool(use_xyz)), ptr(grad_out), ptr(idx), ptr(grad_features), stream()),
              "group_concat_grad")


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    require("group_points_grad", (grad_out_tensor, F32, b * c * npoints * nsample), (idx_tensor, I32, b * npoints * nsample),
            (grad_points_tensor, F32, b * c * n))
    with device_of(grad_out_tensor):
        check(lib().ws3d_group_points_grad(b, c, n, npoints, nsample, ptr(grad_out_tensor), ptr(idx_tensor),
                                           ptr(grad_points_tensor), stream()), "group_points_grad")
    return 1


def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    require("three_nn", (unknown_tensor, F32, b * n * 3), (known_tensor, F32, b * m * 3), (dist2_tensor, F32, b * n * 3),
            (idx_tensor, I32, b * n * 3))
    with device_of(unknown_tensor):
        check(lib().ws3d_three_nn(b, n, m, ptr(unknown_tensor), ptr(known_tensor), ptr(dist2_tensor),
                                  ptr(idx_tensor), stream()), "three_nn")


def three_nn_weights(b, n, m, unknown, known, dist2, idx, weight):
    """Extension: three_nn + the normalised inverse-distance weights of pointnet2_modules.py:139-144 in one launch.
    dist2 (B,n,3) may be None."""
    require("three_nn_weights", (unknown, F32, b * n * 3), (known, F32, b * m * 3), (dist2, F32, b * n * 3), (idx, I32, b * n * 3),
            (weight, F32, b * n * 3))
    with device_of(unknown):
        check(lib().ws3d_three_nn_weights(b, n, m, ptr(unknown), ptr(known), ptr(dist2), ptr(idx), ptr(weight), stream()),
              "three_nn_weights")


def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    require("three_interpolate", (points_tensor, F32, b * c * m), (idx_tensor, I32, b * n * 3), (weight_tensor, F32, b * n * 3),
            (out_tensor, F32, b * c * n))
    with device_of(points_tensor):
        check(lib().ws3d_three_interpolate(b, c, m, n, ptr(points_tensor), ptr(idx_tensor), ptr(weight_tensor),
                                           ptr(out_tensor), stream()), "three_interpolate")


def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    require("three_interpolate_grad", (grad_out_tensor, F32, b * c * n), (idx_tensor, I32, b * n * 3),
            (weight_tensor, F32, b * n * 3), (grad_points_tensor, F32, b * c * m))
    with device_of(grad_out_tensor):
        check(lib().ws3d_three_interpolate_grad(b, c, n, m, ptr(grad_out_tensor), ptr(idx_tensor),
                                                ptr(weight_tensor), ptr(grad_points_tensor), stream()),
              "three_interpolate_grad")


# ---- extensions (no reference counterpart; used by ws3d_b200.pointnet2_utils) -------------------
def furthest_point_sampling_gather(b, n, m, xyz, temp, idx, new_xyz):
    require("furthest_point_sampling_gather", (xyz, F32, b * n * 3), (temp, F32, b * n), (idx, I32, b * m), (new_xyz, F32, b * m * 3))
    with device_of(xyz):
        check(lib().ws3d_furthest_point_sampling_gather(b, n, m, ptr(xyz), ptr(temp), ptr(idx), ptr(new_xyz),
                                                        stream()), "furthest_point_sampling_gather")


def ball_query2(b, n, m, radius0, nsample0, radius1, nsample1, new_xyz, xyz, idx0, idx1):
    require("ball_query2", (new_xyz, F32, b * m * 3), (xyz, F32, b * n * 3), (idx0, I32, b * m * nsample0), (idx1, I32, b * m * nsample1))
    with device_of(xyz):
        check(lib().ws3d_ball_query2(b, n, m, float(radius0), nsample0, float(radius1), nsample1, ptr(new_xyz),
                                     ptr(xyz), ptr(idx0), ptr(idx1), stream()), "ball_query2")


def group_concat(b, n, m, c, nsample, use_xyz, xyz, new_xyz, features, idx, out):
    require("group_concat", (xyz, F32, b * n * 3), (new_xyz, F32, b * m * 3), (features, F32, b * c * n), (idx, I32, b * m * nsample),
            (out, F32, b * (c + (3 if use_xyz else 0)) * m * nsample))
    with device_of(out):
        check(lib().ws3d_group_concat(b, n, m, c, nsample, int(bool(use_xyz)), ptr(xyz), ptr(new_xyz), ptr(features),
                                      ptr(idx), ptr(out), stream()), "group_concat")


def query_and_group(b, n, m, c, radius, nsample, use_xyz, xyz, new_xyz, features, out, idx_out):
    require("query_and_group", (xyz, F32, b * n * 3), (new_xyz, F32, b * m * 3), (features, F32, b * c * n),
            (out, F32, b * (c + (3 if use_xyz else 0)) * m * nsample), (idx_out, I32, b * m * nsample))
    with device_of(out):
        check(lib().ws3d_query_and_group(b, n, m, c, float(radius), nsample, int(bool(use_xyz)), ptr(xyz),
                                         ptr(new_xyz), ptr(features), ptr(out), ptr(idx_out), stream()),
              "query_and_group")


def mlp_layer(b, c_out, c_out_pad, c1, c2, cols, w, shift, x1, x2, out, relu, pool, out_ctot=None, out_coff=0):
    """One BN-folded shared-MLP layer on the tcgen05 tensor cores (see include/ws3d_ops.h).
    `relu`: bit 0 = ReLU, bit 1 = round the output to TF32 (it feeds another layer).
    `out_ctot` / `out_coff`: write channels [out_coff, out_coff + c_out) of an (B, out_ctot, .) tensor."""
    ctot = c_out if out_ctot is None else int(out_ctot)
    require("mlp_layer", (w, F32, c_out_pad * (c1 + c2)), (shift, F32, c_out_pad), (x1, F32, b * c1 * cols), (x2, F32, b * c2 * cols),
            (out, F32, b * ctot * (cols // pool if pool else cols)))
    with device_of(out):
        check(lib().ws3d_mlp_layer_into(b, c_out, c_out_pad, c1, c2, cols, ptr(w), ptr(shift), ptr(x1), ptr(x2), ptr(out),
                                        ctot, int(out_coff), int(relu), int(pool), stream()), "mlp_layer")


# ---- training-mode shared MLP (see include/ws3d_ops.h) -------------------------------------------------
def mlp_layer_stats(b, c_out, c_out_pad, c1, c2, cols, w, zero_shift, x1, x2, y, stats):
    require("mlp_layer_stats", (w, F32, c_out_pad * (c1 + c2)), (zero_shift, F32, c_out_pad), (x1, F32, b * c1 * cols),
            (x2, F32, b * c2 * cols), (y, F32, b * c_out * cols), (stats, F64, 2 * c_out))
    with device_of(y):
        check(lib().ws3d_mlp_layer_stats(b, c_out, c_out_pad, c1, c2, cols, ptr(w), ptr(zero_shift), ptr(x1), ptr(x2), ptr(y), ptr(stats),
                                         stream()), "mlp_layer_stats")


def bn_finalize(c, count, stats, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean, invstd):
    require("bn_finalize", (stats, F64, 2 * c), (gamma, F32, c), (beta, F32, c), (running_mean, F32, c), (running_var, F32, c),
            (scale, F32, c), (shift, F32, c), (mean, F32, c), (invstd, F32, c))
    with device_of(stats):
        check(lib().ws3d_bn_finalize(c, float(count), ptr(stats), ptr(gamma), ptr(beta), float(eps), float(momentum), ptr(running_mean),
                                     ptr(running_var), ptr(scale), ptr(shift), ptr(mean), ptr(invstd), stream()), "bn_finalize")


def bn_relu_apply(b, c, cols, pool, y, scale, shift, flags, z, arg):
    require("bn_relu_apply", (y, F32, b * c * cols), (scale, F32, c), (shift, F32, c), (z, F32, b * c * (cols // pool if pool else cols)),
            (arg, U8, b * c * (cols // pool) if pool else None))
    with device_of(y):
        check(lib().ws3d_bn_relu_apply(b, c, cols, pool, ptr(y), ptr(scale), ptr(shift), int(flags), ptr(z), ptr(arg), stream()),
              "bn_relu_apply")


def bn_relu_bwd_reduce(b, c, cols, pool, y, dz, arg, scale, shift, mean, invstd, flags, sums):
    require("bn_relu_bwd_reduce", (y, F32, b * c * cols), (dz, F32, b * c * (cols // pool if pool else cols)),
            (arg, U8, b * c * (cols // pool) if pool else None), (scale, F32, c), (shift, F32, c), (mean, F32, c), (invstd, F32, c),
            (sums, F64, 2 * c))
    with device_of(y):
        check(lib().ws3d_bn_relu_bwd_reduce(b, c, cols, pool, ptr(y), ptr(dz), ptr(arg), ptr(scale), ptr(shift), ptr(mean), ptr(invstd),
                                            int(flags), ptr(sums), stream()), "bn_relu_bwd_reduce")


def bn_relu_bwd_apply(b, c, cols, pool, y, dz, arg, scale, shift, mean, invstd, flags, sums, count, dy):
    require("bn_relu_bwd_apply", (y, F32, b * c * cols), (dz, F32, b * c * (cols // pool if pool else cols)),
            (arg, U8, b * c * (cols // pool) if pool else None), (scale, F32, c), (shift, F32, c), (mean, F32, c), (invstd, F32, c),
            (sums, F64, 2 * c), (dy, F32, b * c * cols))
    with device_of(y):
        check(lib().ws3d_bn_relu_bwd_apply(b, c, cols, pool, ptr(y), ptr(dz), ptr(arg), ptr(scale), ptr(shift), ptr(mean), ptr(invstd),
                                           int(flags), ptr(sums), float(count), ptr(dy), stream()), "bn_relu_bwd_apply")


def mlp_wgrad(b, c_out, c_in, cols, dy, x, dw, ldw, dw_offset=0):
    """dw[:, dw_offset : dw_offset + c_in] += sum_b dy[b] x[b]^T for a (c_out, ldw) row-major dw."""
    require("mlp_wgrad", (dy, F32, b * c_out * cols), (x, F32, b * c_in * cols), (dw, F32, (c_out - 1) * ldw + dw_offset + c_in))
    from ctypes import c_void_p
    with device_of(dy):
        check(lib().ws3d_mlp_wgrad(b, c_out, c_in, cols, ptr(dy), ptr(x), c_void_p(dw.data_ptr() + 4 * int(dw_offset)), int(ldw), stream()),
              "mlp_wgrad")


def split_pointcloud(pc, xyz, features):
    """Extension: pc (B,N,3+C) -> xyz (B,N,3), features (B,C,N) or None in one launch (lib/net/pointnet2_msg.py:52-60)."""
    b, n, c = pc.size(0), pc.size(1), pc.size(2) - 3
    require("split_pointcloud", (pc, F32, None), (xyz, F32, b * n * 3), (features, F32, b * c * n))
    if pc.dim() != 3 or c < 0 or (c > 0 and features is None):
        raise RuntimeError("split_pointcloud: pc must be (B, N, 3 + C) with a features output when C > 0")
    with device_of(pc):
        check(lib().ws3d_split_pointcloud(b, n, c, ptr(pc), ptr(xyz), ptr(features), stream()), "split_pointcloud")


def pack_rows(xyz, features, rows):
    """Extension: xyz (B,N,3) + features (B,C,N) or None -> rows (B,N,ld) = [xyz | features | zeros], point-major."""
    b, n = xyz.size(0), xyz.size(1)
    c = 0 if features is None else features.size(1)
    ld = rows.size(2)
    require("pack_rows", (xyz, F32, b * n * 3), (features, F32, b * c * n), (rows, F32, b * n * ld))
    if ld < 3 + c:
        raise RuntimeError("pack_rows: rows must hold 3 + C columns")
    with device_of(xyz):
        check(lib().ws3d_pack_rows(b, n, c, ld, ptr(xyz), ptr(features), ptr(rows), stream()), "pack_rows")


# ---- iou3d_cuda -------------------------------------------------------------------------------
def _bev(what, boxes):
    """(N, 5) float32 CUDA BEV boxes, the reference's CHECK_INPUT (iou3d.cpp:10-12) plus dtype / shape."""
    require(what, (boxes, F32, None))
    if boxes.dim() != 2 or boxes.size(1) != 5:
        raise RuntimeError(f"{what}: boxes must be (N, 5) [x1, y1, x2, y2, ry]")


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    _bev("boxes_overlap_bev_gpu", boxes_a)
    _bev("boxes_overlap_bev_gpu", boxes_b)
    require("boxes_overlap_bev_gpu", (ans_overlap, F32, boxes_a.size(0) * boxes_b.size(0)))
    with device_of(boxes_a):
        check(lib().ws3d_boxes_overlap_bev(boxes_a.size(0), ptr(boxes_a), boxes_b.size(0), ptr(boxes_b),
                                           ptr(ans_overlap), stream()), "boxes_overlap_bev_gpu")
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    _bev("boxes_iou_bev_gpu", boxes_a)
    _bev("boxes_iou_bev_gpu", boxes_b)
    require("boxes_iou_bev_gpu", (ans_iou, F32, boxes_a.size(0) * boxes_b.size(0)))
    with device_of(boxes_a):
        check(lib().ws3d_boxes_iou_bev(boxes_a.size(0), ptr(boxes_a), boxes_b.size(0), ptr(boxes_b), ptr(ans_iou),
                                       stream()), "boxes_iou_bev_gpu")
    return 1


def _nms_host(fn, name, boxes, keep, thresh):
    _bev(name, boxes)
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous() or keep.numel() < boxes.size(0):
        raise RuntimeError("keep must be a contiguous CPU int64 tensor with one slot per box")
    with device_of(boxes):
        n = fn(ptr(boxes), boxes.size(0), float(thresh), ptr(keep), stream())
    if n < 0:
        raise RuntimeError(f"ws3d_b200.{name} failed (code {-n}): {_C.last_error()}")
    return n


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """Reference signature (iou3d.cpp:73): boxes (N,5) CUDA sorted by score, keep (N) CPU int64."""
    return _nms_host(lib().ws3d_nms_host, "nms_gpu", boxes, keep, nms_overlap_thresh)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms_host(lib().ws3d_nms_normal_host, "nms_normal_gpu", boxes, keep, nms_overlap_thresh)


def nms_device(boxes, thresh, rotated=True):
    """Extension: all-device NMS.  Returns (keep int64 CUDA (N), num_keep int32 CUDA (1)); no sync."""
    _bev("nms", boxes)
    n = boxes.size(0)
    keep = torch.empty(n, dtype=torch.int64, device=boxes.device)
    num = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    fn = lib().ws3d_nms if rotated else lib().ws3d_nms_normal
    with device_of(boxes):
        check(fn(ptr(boxes), n, float(thresh), ptr(keep), ptr(num), None, stream()), "nms")
    return keep, num


def boxes_iou3d_aligned(boxes_a, boxes_b, iou2d, iou3d):
    """Extension: diagonal of boxes_iou3d_gpu for aligned (n,7) box pairs; outputs (n) each."""
    require("boxes_iou3d_aligned", (boxes_a, F32, None), (boxes_b, F32, None), (iou2d, F32, boxes_a.size(0)), (iou3d, F32, boxes_a.size(0)))
    if boxes_a.shape != boxes_b.shape or boxes_a.dim() != 2 or boxes_a.size(1) != 7:
        raise RuntimeError("boxes_iou3d_aligned: boxes must both be (n, 7)")
    with device_of(boxes_a):
        check(lib().ws3d_boxes_iou3d_aligned(boxes_a.size(0), ptr(boxes_a), ptr(boxes_b), ptr(iou2d), ptr(iou3d), stream()),
              "boxes_iou3d_aligned")


def radius_nms_device(centers, radius):
    """Extension: greedy radius NMS over (n,2) BEV centres sorted by descending score.
    Returns (keep int64 CUDA (n), num_keep int32 CUDA (1)); no sync."""
    require("radius_nms", (centers, F32, None))
    if centers.dim() != 2 or centers.size(1) != 2:
        raise RuntimeError("radius_nms: centers must be (n, 2)")
    n = centers.size(0)
    keep = torch.empty(n, dtype=torch.int64, device=centers.device)
    num = torch.zeros(1, dtype=torch.int32, device=centers.device)
    with device_of(centers):
        check(lib().ws3d_radius_nms(ptr(centers), n, float(radius), ptr(keep), ptr(num), None, stream()), "radius_nms")
    return keep, num


def cylinder_query(pts, centers, radius, idx, cnt, any_flag=None):
    """Extension: pts (n,3), centers (m,2) -> idx (m,cap) int32, cnt (m) int32, any_flag (n) uint8."""
    require("cylinder_query", (pts, F32, None), (centers, F32, None))
    if pts.dim() != 2 or pts.size(1) != 3 or centers.dim() != 2 or centers.size(1) != 2 or idx.dim() != 2:
        raise RuntimeError("cylinder_query: pts must be (n, 3), centers (m, 2) and idx (m, cap)")
    require("cylinder_query", (idx, I32, centers.size(0) * idx.size(1)), (cnt, I32, centers.size(0)), (any_flag, U8, pts.size(0)))
    with device_of(pts):
        check(lib().ws3d_cylinder_query(pts.size(0), centers.size(0), idx.size(1), float(radius), ptr(pts), ptr(centers),
                                        ptr(idx), ptr(cnt), ptr(any_flag), stream()), "cylinder_query")


def gaussian_rpn_labels(pts, gt_boxes3d, num_gt, gauss_height, gauss_status, gauss_cov, fg_radius, cls_label, reg_label):
    """Extension: pts (B,n,3), gt_boxes3d (B,G,7), num_gt (B) int32 or None -> cls_label (B,n), reg_label (B,n,3)."""
    require("gaussian_rpn_labels", (pts, F32, None), (gt_boxes3d, F32, None))
    if pts.dim() != 3 or pts.size(2) != 3 or gt_boxes3d.dim() != 3 or gt_boxes3d.size(2) != 7 or gt_boxes3d.size(0) != pts.size(0):
        raise RuntimeError("gaussian_rpn_labels: pts must be (B, n, 3) and gt_boxes3d (B, G, 7)")
    require("gaussian_rpn_labels", (num_gt, I32, pts.size(0)), (cls_label, F32, pts.size(0) * pts.size(1)),
            (reg_label, F32, pts.size(0) * pts.size(1) * 3))
    with device_of(pts):
        check(lib().ws3d_gaussian_rpn_labels(pts.size(0), pts.size(1), gt_boxes3d.size(1), ptr(pts), ptr(gt_boxes3d), ptr(num_gt),
                                             float(gauss_height), float(gauss_status), float(gauss_cov), float(fg_radius),
                                             ptr(cls_label), ptr(reg_label), stream()), "gaussian_rpn_labels")


def boxes3d_to_corners3d(boxes3d, flip, corners):
    """Extension: boxes3d (N,7) -> corners (N,8,3), kitti_utils.py:104-131."""
    require("boxes3d_to_corners3d", (boxes3d, F32, None), (corners, F32, boxes3d.size(0) * 24))
    if boxes3d.dim() != 2 or boxes3d.size(1) != 7:
        raise RuntimeError("boxes3d_to_corners3d: boxes3d must be (N, 7)")
    with device_of(boxes3d):
        check(lib().ws3d_boxes3d_to_corners3d(boxes3d.size(0), ptr(boxes3d), int(bool(flip)), ptr(corners), stream()),
              "boxes3d_to_corners3d")


def corner_distance(pred, gt, dist):
    """Extension: aligned (N,7) box pairs -> dist (N,8), train_functions.py:266-271."""
    require("corner_distance", (pred, F32, None), (gt, F32, None), (dist, F32, pred.size(0) * 8))
    if pred.dim() != 2 or pred.size(1) != 7 or pred.shape != gt.shape:
        raise RuntimeError("corner_distance: boxes must both be (N, 7)")
    with device_of(pred):
        check(lib().ws3d_corner_distance(pred.size(0), ptr(pred), ptr(gt), ptr(dist), stream()), "corner_distance")


def corner_distance_grad(pred, gt, grad_dist, grad_pred):
    require("corner_distance_grad", (pred, F32, None), (gt, F32, None), (grad_dist, F32, pred.size(0) * 8),
            (grad_pred, F32, pred.size(0) * 7))
    if pred.dim() != 2 or pred.size(1) != 7 or pred.shape != gt.shape:
        raise RuntimeError("corner_distance_grad: boxes must both be (N, 7)")
    with device_of(pred):
        check(lib().ws3d_corner_distance_grad(pred.size(0), ptr(pred), ptr(gt), ptr(grad_dist), ptr(grad_pred), stream()),
              "corner_distance_grad")


def subsample_points(pts, depth, npoints, n_near, near_depth, sub_last, perm, order, out, choice, status):
    """Extension: kitti_rcnn_dataset.py:424-452 on the device with the host's random draws (ws3d_ops.h)."""
    n, c = pts.size(0), pts.size(1)
    need = n_near if n > npoints else npoints
    require("subsample_points", (pts, F32, None), (depth, F32, n), (perm, I32, min(need, npoints) if n <= npoints else npoints - (n - n_near)),
            (order, I32, npoints), (out, F32, npoints * c), (choice, I32, npoints), (status, I32, 1))
    if pts.dim() != 2 or c < 3:
        raise RuntimeError("subsample_points: pts must be (n, 3 + C)")
    with device_of(pts):
        check(lib().ws3d_subsample_points(n, c, int(npoints), int(n_near), float(near_depth), float(sub_last), ptr(pts), ptr(depth),
                                          ptr(perm), ptr(order), ptr(out), ptr(choice), ptr(status), stream()), "subsample_points")


# ---- roipool3d_cuda ---------------------------------------------------------------------------
def roipool3d_forward(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag):
    """Reference `forward` (roipool3d.cpp:48): xyz (B,N,3), boxes3d (B,M,7), pts_feature (B,N,C),
    pooled_features (B,M,S,3+C) and pooled_empty_flag (B,M) int32 zero-filled by the caller."""
    require("roipool3d forward", (xyz, F32, None), (boxes3d, F32, None), (pts_feature, F32, None), (pooled_features, F32, None),
            (pooled_empty_flag, I32, None))
    if (xyz.dim() != 3 or xyz.size(2) != 3 or boxes3d.dim() != 3 or boxes3d.size(2) != 7 or pts_feature.dim() != 3
            or pooled_features.dim() != 4 or boxes3d.size(0) != xyz.size(0) or pts_feature.shape[:2] != xyz.shape[:2]
            or pooled_features.shape[:2] != boxes3d.shape[:2] or pooled_features.size(3) != 3 + pts_feature.size(2)
            or pooled_empty_flag.numel() < boxes3d.size(0) * boxes3d.size(1)):
        raise RuntimeError("roipool3d forward: expected xyz (B,N,3), boxes3d (B,M,7), pts_feature (B,N,C), "
                           "pooled_features (B,M,S,3+C), pooled_empty_flag (B,M)")
    with device_of(xyz):
        check(lib().ws3d_roipool3d(xyz.size(0), xyz.size(1), boxes3d.size(1), pts_feature.size(2),
                                   pooled_features.size(2), ptr(xyz), ptr(boxes3d), ptr(pts_feature),
                                   ptr(pooled_features), ptr(pooled_empty_flag), stream()), "roipool3d forward")
    return 1


def pts_in_boxes3d_cpu(pts_flag, pts, boxes3d):
    for t in (pts_flag, pts, boxes3d):
        if t.is_cuda or not t.is_contiguous():
            raise RuntimeError("pts_in_boxes3d_cpu expects contiguous CPU tensors")
    check(lib().ws3d_pts_in_boxes3d_cpu(ptr(pts_flag), ptr(pts), ptr(boxes3d), boxes3d.size(0), pts.size(0)),
          "pts_in_boxes3d_cpu")
    return 1


def roipool3d_cpu(pts, boxes3d, pts_feature, pooled_pts, pooled_features, pooled_empty_flag):
    for t in (pts, boxes3d, pts_feature, pooled_pts, pooled_features, pooled_empty_flag):
        if t.is_cuda or not t.is_contiguous():
            raise RuntimeError("roipool3d_cpu expects contiguous CPU tensors")
    check(lib().ws3d_roipool3d_cpu(ptr(pts), ptr(boxes3d), ptr(pts_feature), ptr(pooled_pts), ptr(pooled_features),
                                   ptr(pooled_empty_flag), boxes3d.size(0), pts.size(0), pts_feature.size(1),
                                   pooled_pts.size(1)), "roipool3d_cpu")
    return 1


def sa_mlp_fused_supported(c_feat, nsample, c1, c2, c3) -> bool:
    return bool(lib().ws3d_sa_mlp_fused_supported(int(c_feat), int(nsample), int(c1), int(c2), int(c3)))


def sa_mlp_fused(b, n, m, nsample, c_feat, xyz, new_xyz, features, idx, widths, w, shift, out, out_ctot, out_coff):
    """Extension: grouping + 3 x (conv1x1 + BN + ReLU) + max-pool of one set-abstraction scale."""
    require("sa_mlp_fused", (xyz, F32, b * n * 3), (new_xyz, F32, b * m * 3), (features, F32, b * c_feat * n), (idx, I32, b * m * nsample),
            (out, F32, b * out_ctot * m), *[(t, F32, None) for t in list(w) + list(shift)])
    with device_of(xyz):
        check(lib().ws3d_sa_mlp_fused(b, n, m, nsample, c_feat, ptr(xyz), ptr(new_xyz), ptr(features), ptr(idx),
                                      widths[0], widths[1], widths[2], ptr(w[0]), ptr(shift[0]), ptr(w[1]), ptr(shift[1]),
                                      ptr(w[2]), ptr(shift[2]), ptr(out), out_ctot, out_coff, stream()), "sa_mlp_fused")


def sa_mlp_fused_rows(b, n, m, nsample, c_feat, rows, new_xyz, idx, widths, w, shift, out, out_ctot, out_coff, out_pm=None, pm_xyz=False):
    """Extension: ws3d_sa_mlp_fused gathering from point-major rows (B, n, ld) = [x, y, z, features, zeros]; optionally the
    pooled channels also written point-major into out_pm (B, m, ld_pm) for the next level (include/ws3d_ops.h)."""
    ld = rows.shape[-1]
    ld_pm = 0 if out_pm is None else out_pm.shape[-1]
    require("sa_mlp_fused_rows", (rows, F32, b * n * ld), (new_xyz, F32, b * m * 3), (idx, I32, b * m * nsample), (out, F32, b * out_ctot * m),
            (out_pm, F32, b * m * ld_pm), *[(t, F32, None) for t in list(w) + list(shift)])
    with device_of(rows):
        check(lib().ws3d_sa_mlp_fused_rows(b, n, m, nsample, c_feat, ptr(rows), ld, ptr(new_xyz), ptr(idx), widths[0], widths[1], widths[2],
                                           ptr(w[0]), ptr(shift[0]), ptr(w[1]), ptr(shift[1]), ptr(w[2]), ptr(shift[2]), ptr(out), out_ctot,
                                           out_coff, ptr(out_pm), ld_pm, int(bool(pm_xyz)), stream()), "sa_mlp_fused_rows")


def group_affine(b, n, m, c, nsample, P, xyz, new_xyz, wx, shift, idx, flags, out):
    """Extension: out[b,c,j,s] = act(P[b,c,i] + wx[c] . (xyz[b,i] - new_xyz[b,j]) + shift[c]), i = idx[b,j,s]; flags: 1 ReLU, 2 TF32."""
    require("group_affine", (P, F32, b * c * n), (xyz, F32, b * n * 3), (new_xyz, F32, b * m * 3), (wx, F32, c * 3), (shift, F32, c),
            (idx, I32, b * m * nsample), (out, F32, b * c * m * nsample))
    with device_of(P):
        check(lib().ws3d_group_affine(b, n, m, c, nsample, ptr(P), ptr(xyz), ptr(new_xyz), ptr(wx), ptr(shift), ptr(idx),
                                      int(flags), ptr(out), stream()), "group_affine")


def three_interpolate_affine(b, c, m, n, points, idx, weight, scale1, row1, shift, flags, out):
    """Extension: out = act(three_interpolate(points) + scale1[c] * row1[b, i] + shift[c]); flags: 1 ReLU, 2 TF32 rounding."""
    require("three_interpolate_affine", (points, F32, b * c * m), (idx, I32, b * n * 3), (weight, F32, b * n * 3), (scale1, F32, c),
            (row1, F32, b * n), (shift, F32, c), (out, F32, b * c * n))
    with device_of(points):
        check(lib().ws3d_three_interpolate_affine(b, c, m, n, ptr(points), ptr(idx), ptr(weight), ptr(scale1), ptr(row1),
                                                  ptr(shift), int(flags), ptr(out), stream()), "three_interpolate_affine")


def set_workspace_arena(arena: int) -> int:
    """Scratch arena (0..7) for this thread's subsequent launches; returns the previous one (ws3d_ops.h)."""
    return int(lib().ws3d_set_workspace_arena(int(arena)))


def num_arenas() -> int:
    return int(lib().ws3d_num_arenas())


def scratch_bytes(retired_only: bool = False) -> int:
    """Bytes of cached scratch on the current device (ws3d_ops.h)."""
    return int(lib().ws3d_scratch_bytes(int(bool(retired_only))))


def release_scratch(everything: bool = False) -> None:
    """Frees the retired scratch buffers (or every cached buffer) of the current device after a device synchronise.
    Graphs captured while a freed buffer was live must not be replayed again (ws3d_ops.h)."""
    check(lib().ws3d_release_scratch(int(bool(everything))), "release_scratch")


def set_sm_budget(sms: int) -> int:
    """Upper bound on the SMs a persistent kernel (mlp_layer) spreads over; 0 = all.  Returns the previous value."""
    return int(lib().ws3d_set_sm_budget(int(sms)))


def set_fps_mode(mode: int) -> int:
    """0 auto, 1 throughput (several clouds per SM), 2 latency (clusters) for this thread's FPS launches; returns the previous mode."""
    return int(lib().ws3d_set_fps_mode(int(mode)))


def fps_clouds_per_cta(b: int, n: int) -> int:
    """Clouds per CTA (= per SM) of the throughput sampler for a batch of b clouds of n points under this thread's mode."""
    return int(lib().ws3d_fps_clouds_per_cta(int(b), int(n)))
