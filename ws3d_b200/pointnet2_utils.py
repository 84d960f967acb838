"""Host-side mirror of the reference's op wrappers (pointnet2_lib/pointnet2/pointnet2_utils.py).

Same public names, argument meaning and return values -- furthest_point_sample,
gather_operation, three_nn, three_interpolate, grouping_operation, ball_query, QueryAndGroup,
GroupAll -- on top of libws3d_ops.so, plus the fused entry points the B200 modules use
(sample_and_gather, ball_query_pair, query_and_group).  The autograd boundary is the same as in
the reference: index-producing ops have no gradient, gather / group / interpolate have explicit
backward kernels.
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from . import native


def _f32c(t: torch.Tensor) -> torch.Tensor:
    assert t.is_contiguous(), "tensor must be contiguous (same precondition as the reference wrapper)"
    assert t.dtype == torch.float32
    return t


class _FurthestPointSampling(Function):
    """pointnet2_utils.py:10-36.  xyz (B,N,3) -> idx (B,npoint) int32; no gradient."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        _f32c(xyz)
        B, N, _ = xyz.shape
        idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        native.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample = _FurthestPointSampling.apply


def sample_and_gather(xyz: torch.Tensor, npoint: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """FPS with the sampled coordinates emitted by the same kernel.

    Equals (idx, gather_operation(xyz^T, idx)^T) of pointnet2_modules.py:30-35 without the two
    transposes and the gather launch.  Returns idx (B,npoint) int32, new_xyz (B,npoint,3)."""
    _f32c(xyz)
    B, N, _ = xyz.shape
    idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device)
    with torch.no_grad():
        native.furthest_point_sampling_gather(B, N, npoint, xyz, None, idx, new_xyz)
    return idx, new_xyz


def pack_rows(xyz: torch.Tensor, features: Optional[torch.Tensor]) -> torch.Tensor:
    """xyz (B,N,3), features (B,C,N) -> point-major operand rows (B,N,ld) = [xyz | features | zeros], ld a multiple of 8:
    what the fused set-abstraction kernels gather from (one contiguous row per neighbour).  No gradient (inference path)."""
    _f32c(xyz)
    B, N, _ = xyz.shape
    C = 0 if features is None else features.shape[1]
    rows = torch.empty((B, N, (3 + C + 7) // 8 * 8), dtype=torch.float32, device=xyz.device)
    with torch.no_grad():
        native.pack_rows(xyz, None if features is None else _f32c(features.detach()), rows)
    return rows


class _GatherOperation(Function):
    """pointnet2_utils.py:39-73.  features (B,C,N), idx (B,npoint) -> (B,C,npoint)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        _f32c(features)
        assert idx.is_contiguous()
        B, npoint = idx.shape
        _, C, N = features.shape
        out = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
        native.gather_points_wrapper(B, C, N, npoint, features, idx, out)
        ctx.save_for_backward(idx)
        ctx.dims = (C, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        C, N = ctx.dims
        B, npoint = idx.shape
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        native.gather_points_grad_wrapper(B, C, N, npoint, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = _GatherOperation.apply


class _ThreeNN(Function):
    """pointnet2_utils.py:76-105.  Returns (dist, idx): EUCLIDEAN distances (sqrt of the kernel's
    squared distances, as the reference does at :98) and int32 indices; no gradient."""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor):
        _f32c(unknown)
        _f32c(known)
        B, N, _ = unknown.shape
        m = known.shape[1]
        dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
        native.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = _ThreeNN.apply


def three_nn_weights(unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """three_nn and the interpolation weights of PointnetFPModule.forward (pointnet2_modules.py:139-144:
    `1 / (dist + 1e-8)`, normalised over the three neighbours) from one launch.  Returns (idx (B,n,3) int32,
    weight (B,n,3)); no gradient (the reference's three_nn has none and the weights are built from its output)."""
    _f32c(unknown)
    _f32c(known)
    B, N, _ = unknown.shape
    idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
    weight = torch.empty((B, N, 3), dtype=torch.float32, device=unknown.device)
    with torch.no_grad():
        native.three_nn_weights(B, N, known.shape[1], unknown, known, None, idx, weight)
    return idx, weight


class _ThreeInterpolate(Function):
    """pointnet2_utils.py:108-153.  features (B,C,M), idx/weight (B,n,3) -> (B,C,n)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        _f32c(features)
        _f32c(weight)
        assert idx.is_contiguous()
        B, c, m = features.shape
        n = idx.shape[1]
        out = torch.empty((B, c, n), dtype=torch.float32, device=features.device)
        native.three_interpolate_wrapper(B, c, m, n, features, idx, weight, out)
        ctx.save_for_backward(idx, weight)
        ctx.m = m
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        B, c, n = grad_out.shape
        grad_features = torch.zeros((B, c, ctx.m), dtype=torch.float32, device=grad_out.device)
        native.three_interpolate_grad_wrapper(B, c, n, ctx.m, grad_out.contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = _ThreeInterpolate.apply


class _GroupingOperation(Function):
    """pointnet2_utils.py:156-197.  features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        _f32c(features)
        assert idx.is_contiguous()
        B, npoint, nsample = idx.shape
        _, C, N = features.shape
        out = torch.empty((B, C, npoint, nsample), dtype=torch.float32, device=features.device)
        native.group_points_wrapper(B, C, N, npoint, nsample, features, idx, out)
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, C, npoint, nsample = grad_out.shape
        grad_features = torch.zeros((B, C, ctx.N), dtype=torch.float32, device=grad_out.device)
        native.group_points_grad_wrapper(B, C, ctx.N, npoint, nsample, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


grouping_operation = _GroupingOperation.apply


class _BallQuery(Function):
    """pointnet2_utils.py:200-228.  NOTE the public argument order (radius, nsample, xyz, new_xyz)."""

    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        _f32c(xyz)
        _f32c(new_xyz)
        B, N, _ = xyz.shape
        npoint = new_xyz.shape[1]
        idx = torch.empty((B, npoint, nsample), dtype=torch.int32, device=xyz.device)   # every row is written (ws3d_ops.h)
        native.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = _BallQuery.apply


def ball_query_pair(radii, nsamples, xyz: torch.Tensor, new_xyz: torch.Tensor):
    """Two ball queries over the same centres in one scan (multi-scale grouping).  Each returned
    idx equals ball_query(radius, nsample, xyz, new_xyz) for its scale."""
    _f32c(xyz)
    _f32c(new_xyz)
    B, N, _ = xyz.shape
    npoint = new_xyz.shape[1]
    idx0 = torch.empty((B, npoint, nsamples[0]), dtype=torch.int32, device=xyz.device)   # every row is written (ws3d_ops.h)
    idx1 = torch.empty((B, npoint, nsamples[1]), dtype=torch.int32, device=xyz.device)
    with torch.no_grad():
        native.ball_query2(B, N, npoint, radii[0], nsamples[0], radii[1], nsamples[1], new_xyz, xyz, idx0, idx1)
    return idx0, idx1


class _GroupConcat(Function):
    """Fused grouping of QueryAndGroup.forward (pointnet2_utils.py:250-257) for a given idx:
    [xyz[idx] - centre ; features[idx]] written once.  Gradient flows to `features` only (the
    coordinates carry no gradient in any WS3D model; callers that need d/dxyz use the unfused ops)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, features, idx, use_xyz: bool):
        B, N, _ = xyz.shape
        _, M, K = idx.shape
        C = 0 if features is None else features.shape[1]
        out = torch.empty((B, C + (3 if use_xyz else 0), M, K), dtype=torch.float32, device=xyz.device)
        native.group_concat(B, N, M, C, K, use_xyz, xyz, new_xyz, features, idx, out)
        ctx.save_for_backward(idx)
        ctx.dims = (N, C, 3 if use_xyz else 0)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        N, C, off = ctx.dims
        grad_features = None
        if C > 0 and ctx.needs_input_grad[2]:
            B, _, M, K = grad_out.shape
            grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
            # the coordinate channels are skipped by addressing: no contiguous copy of grad_out[:, 3:]
            native.group_concat_grad(B, N, M, C, K, off == 3, grad_out.contiguous(), idx, grad_features)
        return None, None, grad_features, None, None


def group_concat(xyz, new_xyz, features, idx, use_xyz=True):
    return _GroupConcat.apply(xyz, new_xyz, features, idx, use_xyz)


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:231-264: ball query + grouping; returns (B, 3+C, npoint, nsample)."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: Optional[torch.Tensor] = None,
                idx: Optional[torch.Tensor] = None) -> torch.Tensor:
        if idx is None:
            idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        if xyz.requires_grad or new_xyz.requires_grad:
            # coordinates with gradient: compose from the differentiable primitives like the reference
            grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
            grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
            if features is None:
                return grouped_xyz
            grouped = grouping_operation(features, idx)
            return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
        return group_concat(xyz, new_xyz, features, idx, self.use_xyz)


class GroupAll(nn.Module):
    """pointnet2_utils.py:267-290: every point is one group; returns (B, 3+C, 1, N)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz=None, features: Optional[torch.Tensor] = None, idx=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
