"""ctypes binding of libws3d_ops.so (C ABI declared in include/ws3d_ops.h).

The library is the product: there is NO fallback.  If it is missing or fails to load this
module raises, and every wrapper raises RuntimeError when a launch fails (the reference calls
exit(); see SURVEY.md section 8b "errors").
"""
import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libws3d_ops.so")

_vp, _i, _f = c_void_p, c_int, c_float

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "ws3d_abi_version": [],
    "ws3d_last_error": [],
    "ws3d_launch_count": [],
    "ws3d_set_workspace_arena": [_i],
    "ws3d_num_arenas": [],
    "ws3d_scratch_bytes": [_i],
    "ws3d_release_scratch": [_i],
    "ws3d_set_sm_budget": [_i],
    "ws3d_set_fps_mode": [_i],
    "ws3d_fps_clouds_per_cta": [_i, _i],
    "ws3d_furthest_point_sampling": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "ws3d_furthest_point_sampling_gather": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "ws3d_gather_points": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "ws3d_gather_points_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "ws3d_ball_query": [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp],
    "ws3d_ball_query2": [_i, _i, _i, _f, _i, _f, _i, _vp, _vp, _vp, _vp, _vp],
    "ws3d_group_points": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "ws3d_group_points_grad": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "ws3d_pack_rows": [_i, _i, _i, _i, _vp, _vp, _vp],
    "ws3d_group_concat_grad": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "ws3d_query_and_group": [_i, _i, _i, _i, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "ws3d_group_concat": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "ws3d_group_affine": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp],
    "ws3d_three_nn": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "ws3d_three_nn_weights": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "ws3d_three_interpolate": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "ws3d_three_interpolate_affine": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp],
    "ws3d_three_interpolate_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "ws3d_mlp_layer": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "ws3d_mlp_layer_into": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "ws3d_mlp_layer_stats": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ws3d_bn_finalize": [_i, ctypes.c_double, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ws3d_bn_relu_apply": [_i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp],
    "ws3d_bn_relu_bwd_reduce": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp],
    "ws3d_bn_relu_bwd_apply": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, ctypes.c_double, _vp, _vp],
    "ws3d_mlp_wgrad": [_i, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    "ws3d_split_pointcloud": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "ws3d_sa_mlp_fused_supported": [_i, _i, _i, _i, _i],
    "ws3d_sa_mlp_fused": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "ws3d_sa_mlp_fused_rows": [_i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _vp],
    "ws3d_boxes_overlap_bev": [_i, _vp, _i, _vp, _vp, _vp],
    "ws3d_boxes_iou_bev": [_i, _vp, _i, _vp, _vp, _vp],
    "ws3d_nms_workspace_bytes": [_i],
    "ws3d_nms": [_vp, _i, _f, _vp, _vp, _vp, _vp],
    "ws3d_nms_normal": [_vp, _i, _f, _vp, _vp, _vp, _vp],
    "ws3d_nms_host": [_vp, _i, _f, _vp, _vp],
    "ws3d_nms_normal_host": [_vp, _i, _f, _vp, _vp],
    "ws3d_boxes_iou3d_aligned": [_i, _vp, _vp, _vp, _vp, _vp],
    "ws3d_radius_nms": [_vp, _i, _f, _vp, _vp, _vp, _vp],
    "ws3d_cylinder_query": [_i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp],
    "ws3d_gaussian_rpn_labels": [_i, _i, _i, _vp, _vp, _vp, _f, _f, _f, _f, _vp, _vp, _vp],
    "ws3d_boxes3d_to_corners3d": [_i, _vp, _i, _vp, _vp],
    "ws3d_corner_distance": [_i, _vp, _vp, _vp, _vp],
    "ws3d_corner_distance_grad": [_i, _vp, _vp, _vp, _vp, _vp],
    "ws3d_subsample_points": [_i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ws3d_roipool3d": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "ws3d_pts_in_boxes3d_cpu": [_vp, _vp, _vp, _i, _i],
    "ws3d_roipool3d_cpu": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i],
}
_RESTYPES = {
    "ws3d_last_error": ctypes.c_char_p,
    "ws3d_launch_count": c_uint64,
    "ws3d_nms_workspace_bytes": c_size_t,
    "ws3d_scratch_bytes": c_size_t,
}
EXPORTS = tuple(sorted(_SIGNATURES))

_lib = None


def lib() -> ctypes.CDLL:
    """The loaded library (loads it on first use; raises if it is not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m ws3d_b200.build` "
                "(nvcc, sm_100a).  ws3d_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here == ABI mismatch: fail loudly
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        if handle.ws3d_abi_version() != 1:
            raise ImportError("libws3d_ops.so ABI version mismatch")
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().ws3d_last_error().decode("utf-8", "replace")


def launch_count() -> int:
    return int(lib().ws3d_launch_count())


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"ws3d_b200.{what} failed (code {rc}): {last_error()}")


def ptr(t):
    """Device (or host) address of a tensor's first element; None -> NULL."""
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    """torch's current CUDA stream as a cudaStream_t (the reference's pointnet2 ops use the same)."""
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors, what: str = "ws3d_b200"):
    """Same preconditions the reference wrappers assert (CUDA, contiguous) for tensors whose dtype / size are checked
    elsewhere."""
    for k, t in enumerate(tensors):
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"{what}: argument {k} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{what}: argument {k} must be contiguous")
    return True


F32, F64, I32, I64, U8 = torch.float32, torch.float64, torch.int32, torch.int64, torch.uint8


def require(what: str, *specs, cuda: bool = True):
    """Argument validation of one launch.  Every spec is (tensor or None, dtype, minimum element count or None).

    The C ABI takes raw pointers, so a tensor of the wrong dtype would be silently reinterpreted (int64 indices from
    argsort / topk read as int32 pairs, float64 features read as float32) and one that is smaller than the dimensions
    passed beside it would be read or written out of bounds.  The reference's pybind layer raises for the first
    (`tensor.data<int>()` / `data<float>()` type-check) and checks nothing for the second; here both raise RuntimeError."""
    for k, (t, dtype, numel) in enumerate(specs):
        if t is None:
            continue
        if cuda and not t.is_cuda:
            raise RuntimeError(f"{what}: argument {k} must be a CUDA tensor")
        if not cuda and t.is_cuda:
            raise RuntimeError(f"{what}: argument {k} must be a CPU tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{what}: argument {k} must be contiguous")
        if t.dtype != dtype:
            raise RuntimeError(f"{what}: argument {k} must be {dtype}, got {t.dtype}")
        if numel is not None and t.numel() < int(numel):
            raise RuntimeError(f"{what}: argument {k} has {t.numel()} elements, the dimensions passed need {int(numel)}")
    return True


class device_of:
    """Make the tensor's device current for the duration of a launch (the reference has no guard
    and silently requires it; DDP ranks call torch.cuda.set_device so this is normally a no-op)."""

    __slots__ = ("idx", "prev")

    def __init__(self, t):
        self.idx = t.device.index if t.is_cuda else None
        self.prev = None

    def __enter__(self):
        if self.idx is not None:
            cur = torch.cuda.current_device()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False
