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
from ._C import check, device_of, lib, ptr, require_cuda, stream


# ---- pointnet2_cuda ---------------------------------------------------------------------------
def furthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    require_cuda(points_tensor, temp_tensor, idx_tensor)
    with device_of(points_tensor):
        check(lib().ws3d_furthest_point_sampling(b, n, m, ptr(points_tensor), ptr(temp_tensor), ptr(idx_tensor),
                                                 stream()), "furthest_point_sampling")
    return 1


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    require_cuda(points_tensor, idx_tensor, out_tensor)
    with device_of(points_tensor):
        check(lib().ws3d_gather_points(b, c, n, npoints, ptr(points_tensor), ptr(idx_tensor), ptr(out_tensor),
                                       stream()), "gather_points")
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    require_cuda(grad_out_tensor, idx_tensor, grad_points_tensor)
    with device_of(grad_out_tensor):
        check(lib().ws3d_gather_points_grad(b, c, n, npoints, ptr(grad_out_tensor), ptr(idx_tensor),
                                            ptr(grad_points_tensor), stream()), "gather_points_grad")
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    require_cuda(new_xyz_tensor, xyz_tensor, idx_tensor)
    with device_of(xyz_tensor):
        check(lib().ws3d_ball_query(b, n, m, float(radius), nsample, ptr(new_xyz_tensor), ptr(xyz_tensor),
                                    ptr(idx_tensor), stream()), "ball_query")
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    require_cuda(points_tensor, idx_tensor, out_tensor)
    with device_of(points_tensor):
        check(lib().ws3d_group_points(b, c, n, npoints, nsample, ptr(points_tensor), ptr(idx_tensor),
                                      ptr(out_tensor), stream()), "group_points")
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    require_cuda(grad_out_tensor, idx_tensor, grad_points_tensor)
    with device_of(grad_out_tensor):
        check(lib().ws3d_group_points_grad(b, c, n, npoints, nsample, ptr(grad_out_tensor), ptr(idx_tensor),
                                           ptr(grad_points_tensor), stream()), "group_points_grad")
    return 1


def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    require_cuda(unknown_tensor, known_tensor, dist2_tensor, idx_tensor)
    with device_of(unknown_tensor):
        check(lib().ws3d_three_nn(b, n, m, ptr(unknown_tensor), ptr(known_tensor), ptr(dist2_tensor),
                                  ptr(idx_tensor), stream()), "three_nn")


def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    require_cuda(points_tensor, idx_tensor, weight_tensor, out_tensor)
    with device_of(points_tensor):
        check(lib().ws3d_three_interpolate(b, c, m, n, ptr(points_tensor), ptr(idx_tensor), ptr(weight_tensor),
                                           ptr(out_tensor), stream()), "three_interpolate")


def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    require_cuda(grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor)
    with device_of(grad_out_tensor):
        check(lib().ws3d_three_interpolate_grad(b, c, n, m, ptr(grad_out_tensor), ptr(idx_tensor),
                                                ptr(weight_tensor), ptr(grad_points_tensor), stream()),
              "three_interpolate_grad")


# ---- extensions (no reference counterpart; used by ws3d_b200.pointnet2_utils) -------------------
def furthest_point_sampling_gather(b, n, m, xyz, temp, idx, new_xyz):
    require_cuda(xyz, temp, idx, new_xyz)
    with device_of(xyz):
        check(lib().ws3d_furthest_point_sampling_gather(b, n, m, ptr(xyz), ptr(temp), ptr(idx), ptr(new_xyz),
                                                        stream()), "furthest_point_sampling_gather")


def ball_query2(b, n, m, radius0, nsample0, radius1, nsample1, new_xyz, xyz, idx0, idx1):
    require_cuda(new_xyz, xyz, idx0, idx1)
    with device_of(xyz):
        check(lib().ws3d_ball_query2(b, n, m, float(radius0), nsample0, float(radius1), nsample1, ptr(new_xyz),
                                     ptr(xyz), ptr(idx0), ptr(idx1), stream()), "ball_query2")


def group_concat(b, n, m, c, nsample, use_xyz, xyz, new_xyz, features, idx, out):
    require_cuda(xyz, new_xyz, features, idx, out)
    with device_of(out):
        check(lib().ws3d_group_concat(b, n, m, c, nsample, int(bool(use_xyz)), ptr(xyz), ptr(new_xyz), ptr(features),
                                      ptr(idx), ptr(out), stream()), "group_concat")


def query_and_group(b, n, m, c, radius, nsample, use_xyz, xyz, new_xyz, features, out, idx_out):
    require_cuda(xyz, new_xyz, features, out, idx_out)
    with device_of(out):
        check(lib().ws3d_query_and_group(b, n, m, c, float(radius), nsample, int(bool(use_xyz)), ptr(xyz),
                                         ptr(new_xyz), ptr(features), ptr(out), ptr(idx_out), stream()),
              "query_and_group")


def mlp_layer(b, c_out, c_out_pad, c1, c2, cols, w, shift, x1, x2, out, relu, pool):
    """One BN-folded shared-MLP layer on the tcgen05 tensor cores (see include/ws3d_ops.h).
    `relu`: bit 0 = ReLU, bit 1 = round the output to TF32 (it feeds another layer)."""
    require_cuda(w, shift, x1, x2, out)
    with device_of(out):
        check(lib().ws3d_mlp_layer(b, c_out, c_out_pad, c1, c2, cols, ptr(w), ptr(shift), ptr(x1), ptr(x2), ptr(out),
                                   int(relu), int(pool), stream()), "mlp_layer")


# ---- iou3d_cuda -------------------------------------------------------------------------------
def _check_boxes(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("boxes must be a CUDAtensor ")
        if not t.is_contiguous():
            raise RuntimeError("boxes must be contiguous ")


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    _check_boxes(boxes_a, boxes_b, ans_overlap)
    with device_of(boxes_a):
        check(lib().ws3d_boxes_overlap_bev(boxes_a.size(0), ptr(boxes_a), boxes_b.size(0), ptr(boxes_b),
                                           ptr(ans_overlap), stream()), "boxes_overlap_bev_gpu")
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    _check_boxes(boxes_a, boxes_b, ans_iou)
    with device_of(boxes_a):
        check(lib().ws3d_boxes_iou_bev(boxes_a.size(0), ptr(boxes_a), boxes_b.size(0), ptr(boxes_b), ptr(ans_iou),
                                       stream()), "boxes_iou_bev_gpu")
    return 1


def _nms_host(fn, name, boxes, keep, thresh):
    _check_boxes(boxes)
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous():
        raise RuntimeError("keep must be a contiguous CPU int64 tensor")
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
    _check_boxes(boxes)
    n = boxes.size(0)
    keep = torch.empty(n, dtype=torch.int64, device=boxes.device)
    num = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    fn = lib().ws3d_nms if rotated else lib().ws3d_nms_normal
    with device_of(boxes):
        check(fn(ptr(boxes), n, float(thresh), ptr(keep), ptr(num), None, stream()), "nms")
    return keep, num


def boxes_iou3d_aligned(boxes_a, boxes_b, iou2d, iou3d):
    """Extension: diagonal of boxes_iou3d_gpu for aligned (n,7) box pairs; outputs (n) each."""
    _check_boxes(boxes_a, boxes_b, iou2d, iou3d)
    if boxes_a.shape != boxes_b.shape or boxes_a.dim() != 2 or boxes_a.size(1) != 7:
        raise RuntimeError("boxes_iou3d_aligned: boxes must both be (n, 7)")
    with device_of(boxes_a):
        check(lib().ws3d_boxes_iou3d_aligned(boxes_a.size(0), ptr(boxes_a), ptr(boxes_b), ptr(iou2d), ptr(iou3d), stream()),
              "boxes_iou3d_aligned")


def radius_nms_device(centers, radius):
    """Extension: greedy radius NMS over (n,2) BEV centres sorted by descending score.
    Returns (keep int64 CUDA (n), num_keep int32 CUDA (1)); no sync."""
    _check_boxes(centers)
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
    _check_boxes(pts, centers, idx, cnt, any_flag)
    if pts.dim() != 2 or pts.size(1) != 3 or centers.dim() != 2 or centers.size(1) != 2:
        raise RuntimeError("cylinder_query: pts must be (n, 3) and centers (m, 2)")
    with device_of(pts):
        check(lib().ws3d_cylinder_query(pts.size(0), centers.size(0), idx.size(1), float(radius), ptr(pts), ptr(centers),
                                        ptr(idx), ptr(cnt), ptr(any_flag), stream()), "cylinder_query")


def gaussian_rpn_labels(pts, gt_boxes3d, num_gt, gauss_height, gauss_status, gauss_cov, fg_radius, cls_label, reg_label):
    """Extension: pts (B,n,3), gt_boxes3d (B,G,7), num_gt (B) int32 or None -> cls_label (B,n), reg_label (B,n,3)."""
    _check_boxes(pts, gt_boxes3d, num_gt, cls_label, reg_label)
    if pts.dim() != 3 or pts.size(2) != 3 or gt_boxes3d.dim() != 3 or gt_boxes3d.size(2) != 7 or gt_boxes3d.size(0) != pts.size(0):
        raise RuntimeError("gaussian_rpn_labels: pts must be (B, n, 3) and gt_boxes3d (B, G, 7)")
    with device_of(pts):
        check(lib().ws3d_gaussian_rpn_labels(pts.size(0), pts.size(1), gt_boxes3d.size(1), ptr(pts), ptr(gt_boxes3d), ptr(num_gt),
                                             float(gauss_height), float(gauss_status), float(gauss_cov), float(fg_radius),
                                             ptr(cls_label), ptr(reg_label), stream()), "gaussian_rpn_labels")


# ---- roipool3d_cuda ---------------------------------------------------------------------------
def roipool3d_forward(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag):
    """Reference `forward` (roipool3d.cpp:48): xyz (B,N,3), boxes3d (B,M,7), pts_feature (B,N,C),
    pooled_features (B,M,S,3+C) and pooled_empty_flag (B,M) int32 zero-filled by the caller."""
    _check_boxes(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag)
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
    require_cuda(xyz, new_xyz, features, idx, out, *w, *shift)
    with device_of(xyz):
        check(lib().ws3d_sa_mlp_fused(b, n, m, nsample, c_feat, ptr(xyz), ptr(new_xyz), ptr(features), ptr(idx),
                                      widths[0], widths[1], widths[2], ptr(w[0]), ptr(shift[0]), ptr(w[1]), ptr(shift[1]),
                                      ptr(w[2]), ptr(shift[2]), ptr(out), out_ctot, out_coff, stream()), "sa_mlp_fused")


def group_affine(b, n, m, c, nsample, P, xyz, new_xyz, wx, shift, idx, flags, out):
    """Extension: out[b,c,j,s] = act(P[b,c,i] + wx[c] . (xyz[b,i] - new_xyz[b,j]) + shift[c]), i = idx[b,j,s]; flags: 1 ReLU, 2 TF32."""
    require_cuda(P, xyz, new_xyz, wx, shift, idx, out)
    with device_of(P):
        check(lib().ws3d_group_affine(b, n, m, c, nsample, ptr(P), ptr(xyz), ptr(new_xyz), ptr(wx), ptr(shift), ptr(idx),
                                      int(flags), ptr(out), stream()), "group_affine")


def three_interpolate_affine(b, c, m, n, points, idx, weight, scale1, row1, shift, flags, out):
    """Extension: out = act(three_interpolate(points) + scale1[c] * row1[b, i] + shift[c]); flags: 1 ReLU, 2 TF32 rounding."""
    require_cuda(points, idx, weight, scale1, row1, shift, out)
    with device_of(points):
        check(lib().ws3d_three_interpolate_affine(b, c, m, n, ptr(points), ptr(idx), ptr(weight), ptr(scale1), ptr(row1),
                                                  ptr(shift), int(flags), ptr(out), stream()), "three_interpolate_affine")


def set_workspace_arena(arena: int) -> int:
    """Scratch arena (0..7) for this thread's subsequent launches; returns the previous one (ws3d_ops.h)."""
    return int(lib().ws3d_set_workspace_arena(int(arena)))


def set_sm_budget(sms: int) -> int:
    """Upper bound on the SMs a persistent kernel (mlp_layer) spreads over; 0 = all.  Returns the previous value."""
    return int(lib().ws3d_set_sm_budget(int(sms)))


def set_fps_mode(mode: int) -> int:
    """0 auto, 1 throughput (one SM per cloud), 2 latency (clusters) for this thread's FPS launches; returns the previous mode."""
    return int(lib().ws3d_set_fps_mode(int(mode)))
