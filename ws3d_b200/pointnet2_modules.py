"""Host-side mirror of the reference's set-abstraction / feature-propagation modules
(pointnet2_lib/pointnet2/pointnet2_modules.py): PointnetSAModuleMSG, PointnetSAModule and
PointnetFPModule with the same constructor arguments, sub-module names (`groupers`, `mlps`, `mlp`)
and therefore the same state_dict keys.  The forward passes sequence the B200 ops:

  SA layer (reference :19-55)                 here
  ------------------------------------------  ------------------------------------------------
  transpose + FPS + gather + transpose        one FPS launch that also emits new_xyz
  per scale: ball_query                       one scan for both radii (ball_query_pair)
  per scale: group xyz, subtract, group       one grouping pass writing the concatenated
             features, cat                    (3+C, npoint, nsample) tensor once
  SharedMLP + max-pool                        unchanged (PyTorch / cuDNN)
"""
import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused_mlp
from . import pointnet2_utils
from . import pytorch_utils as pt_utils
from . import train_mlp


def _train_channels_last(x: torch.Tensor) -> bool:
    """Training-mode shared MLPs run on channels-last activations (WS3D_TRAIN_CHANNELS_LAST=0 keeps NCHW).
    Measured on B200 (tools/train_profile.py): on the NCHW (B, C, npoint, nsample) tensors cuDNN's training BatchNorm
    takes `bn_fw_tr_1C11` / `bn_bw_1C11`, one thread block per channel -- 31 of the 55 ms of a Stage-1 RPN training step
    (ATen's native kernels are slower still: 41 ms); channels-last selects the NHWC semi-persistent kernels (6 ms) and
    the 1x1 convolutions become plain GEMMs: 55 -> 31 ms per step."""
    return x.is_cuda and torch.is_grad_enabled() and os.environ.get("WS3D_TRAIN_CHANNELS_LAST", "1") != "0"


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.pool_method = 'max_pool'

    def _folded(self, k):
        cache = self.__dict__.setdefault("_folded_mlps", {})
        if k not in cache:
            cache[k] = fused_mlp.FoldedMLP(self.mlps[k])
        return cache[k]

    def _scale_inference(self, k, grouper, xyz, new_xyz, features, idx, out=None, out_coff=0):
        """One scale on the per-layer tensor-core path (inference): (B, c_last, npoint), or its slot of `out`."""
        K = grouper.nsample
        if (features is not None and getattr(grouper, "use_xyz", False) and len(self.mlps[k]) >= 2 and features.is_contiguous()
                and os.environ.get("WS3D_SA_PREMUL", "1") != "0"
                and fused_mlp.FoldedSAFirstLayer.eligible(self.mlps[k][0], features.shape[1], features.shape[2])):
            # first convolution on the source points, gathered with the coordinate term / shift / ReLU in the epilogue
            cache = self.__dict__.setdefault("_premul", {})
            if k not in cache:
                cache[k] = (fused_mlp.FoldedSAFirstLayer(self.mlps[k][0], features.shape[1]),
                            fused_mlp.FoldedMLP(torch.nn.Sequential(*list(self.mlps[k])[1:])))
            first, rest = cache[k]
            return rest(first(xyz, new_xyz, features, idx, round_out=True), pool=K, out=out, out_coff=out_coff)
        grouped = grouper(xyz, new_xyz, features, idx=idx)  # (B, 3+C, npoint, nsample)
        B, C, M, _ = grouped.shape
        return self._folded(k)(grouped.view(B, C, M * K), pool=K, out=out, out_coff=out_coff)

    def _scale_side_streams(self, xyz):
        """One extra CUDA stream per scale beyond the first (cached per module and device)."""
        if not xyz.is_cuda or len(self.groupers) < 2:
            return None
        side = self.__dict__.get("_scale_streams")
        if side is None or side[0].device != xyz.device:
            side = self.__dict__["_scale_streams"] = [torch.cuda.Stream(device=xyz.device) for _ in self.groupers[1:]]
        return side

    def _neighbour_indices(self, xyz, new_xyz):
        """ball_query for every scale; scales are scanned two at a time."""
        specs = [(g.radius, g.nsample) for g in self.groupers]
        out = [None] * len(specs)
        k = 0
        while k + 1 < len(specs):
            (r0, s0), (r1, s1) = specs[k], specs[k + 1]
            out[k], out[k + 1] = pointnet2_utils.ball_query_pair((r0, r1), (s0, s1), xyz, new_xyz)
            k += 2
        if k < len(specs):
            out[k] = pointnet2_utils.ball_query(specs[k][0], specs[k][1], xyz, new_xyz)
        return out

    def fused_scales_eligible(self, c_feat: int) -> bool:
        """True when every scale of this layer runs as one fused kernel in inference (grouping + MLP + max-pool)."""
        return (self.npoint is not None and self.pool_method == 'max_pool' and os.environ.get("WS3D_SA_FUSED", "1") != "0"
                and all(getattr(g, "use_xyz", False) for g in self.groupers)
                and all(fused_mlp.FusedSAScale.eligible(mlp, c_feat, g.nsample) for g, mlp in zip(self.groupers, self.mlps)))

    def forward(self, xyz: torch.Tensor, features: Optional[torch.Tensor] = None,
                new_xyz: Optional[torch.Tensor] = None, indices=None, rows: Optional[torch.Tensor] = None,
                want_rows: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """xyz (B,N,3), features (B,C,N) -> new_xyz (B,npoint,3), new_features (B, sum_k mlps[k][-1], npoint).
        `new_xyz` / `indices` (one ball-query result per scale): the coordinate-only part of the layer if the caller
        has already computed it (models.Pointnet2MSG.coordinate_phase).
        `rows` (B,N,ld) = [xyz | features | zeros], point-major: what the fused kernels gather from when given (inference);
        `want_rows`: also return the layer's output in that layout -- (new_xyz, new_features, rows_out or None)."""
        if new_xyz is None and self.npoint is not None:
            _, new_xyz = pointnet2_utils.sample_and_gather(xyz, self.npoint)
        if self.npoint is not None and indices is not None:
            assert len(indices) == len(self.groupers)
        elif self.npoint is not None:
            indices = self._neighbour_indices(xyz, new_xyz)
        else:
            indices = [None] * len(self.groupers)  # GroupAll
        pooled = []
        fused = fused_mlp.enabled_for(self) and self.pool_method == 'max_pool'
        if fused and self.npoint is not None and os.environ.get("WS3D_SA_FUSED", "1") != "0":
            # inference: one kernel per scale (grouping + 3 layers + max-pool, activations stay in tensor memory),
            # each writing its slot of the concatenated output
            c_feat = 0 if features is None else features.shape[1]
            if (all(getattr(g, "use_xyz", False) for g in self.groupers) and xyz.is_contiguous()
                    and (features is None or features.is_contiguous())
                    and all(fused_mlp.FusedSAScale.eligible(mlp, c_feat, g.nsample) for g, mlp in zip(self.groupers, self.mlps))):
                cache = self.__dict__.setdefault("_fused_scales", {})
                widths = [mlp[-1].conv.out_channels for mlp in self.mlps]
                out = torch.empty((xyz.shape[0], sum(widths), new_xyz.shape[1]), dtype=torch.float32, device=xyz.device)
                if rows is None and c_feat >= 4 and os.environ.get("WS3D_SA_ROWS", "1") != "0":
                    # channel-major features only: one transposing pass builds the point-major operand rows; every point is
                    # gathered npoint * nsample / N times (8 - 12 at the WS3D shapes), each time as ONE contiguous row instead
                    # of c_feat separate 32-byte sectors (Stage-2 level 1: 8.8 GB -> 1.1 GB of L2 traffic per launch)
                    rows = pointnet2_utils.pack_rows(xyz, features)
                if rows is not None and not (rows.is_contiguous() and rows.shape[-1] % 4 == 0 and rows.shape[-1] >= 3 + c_feat
                                             and rows.data_ptr() % 16 == 0 and os.environ.get("WS3D_SA_ROWS", "1") != "0"):
                    rows = None
                out_pm = None
                if want_rows and rows is not None:
                    ld_pm = (3 + sum(widths) + 7) // 8 * 8
                    out_pm = torch.empty((xyz.shape[0], new_xyz.shape[1], ld_pm), dtype=torch.float32, device=xyz.device)
                streams = self._scale_side_streams(xyz) if os.environ.get("WS3D_SCALE_STREAMS", "1") != "0" else None
                main = torch.cuda.current_stream(xyz.device)
                for k, mlp in enumerate(self.mlps):
                    # folded weights are built (small kernels on the CURRENT stream) BEFORE the fork event is recorded: a
                    # side stream that waits for the fork then also waits for its weights.  (Found by compute-sanitizer:
                    # building them after the fork left the first call of a scale on a side stream unordered against it.)
                    if k not in cache:
                        cache[k] = fused_mlp.FusedSAScale(mlp)
                if streams:
                    fork = torch.cuda.Event()
                    fork.record(main)
                off, joins = 0, []
                for k, (mlp, idx) in enumerate(zip(self.mlps, indices)):
                    if streams and k > 0:
                        st = streams[k - 1]
                        st.wait_event(fork)
                        with torch.cuda.stream(st):
                            cache[k](xyz, new_xyz, features, idx, out, off, rows=rows, out_pm=out_pm, pm_xyz=False)
                            done = torch.cuda.Event()
                            done.record(st)
                        for t in (xyz, new_xyz, features, idx, out, rows, out_pm):
                            if t is not None:
                                t.record_stream(st)
                        joins.append(done)
                    else:
                        cache[k](xyz, new_xyz, features, idx, out, off, rows=rows, out_pm=out_pm, pm_xyz=(k == 0))
                    off += widths[k]
                for ev in joins:
                    main.wait_event(ev)
                return (new_xyz, out, out_pm) if want_rows else (new_xyz, out)
        # inference, several scales: the scales are independent chains of small launches (group + three layers, each
        # well under one wave at the deeper levels), so they run side by side on per-scale streams
        side = None
        all_layers = (fused and xyz.is_cuda and self.npoint is not None
                      and all(fused_mlp.supported(new_xyz.shape[1] * g.nsample, g.nsample) for g in self.groupers))
        if all_layers:
            # every scale's last layer writes its slot of the concatenated output (no torch.cat pass)
            widths = [mlp[-1].conv.out_channels for mlp in self.mlps]
            cat_out = torch.empty((xyz.shape[0], sum(widths), new_xyz.shape[1]), dtype=torch.float32, device=xyz.device)
            offs = [sum(widths[:k]) for k in range(len(widths))]
        if all_layers and len(self.groupers) > 1 and os.environ.get("WS3D_SCALE_STREAMS", "1") != "0":
            side = self._scale_side_streams(xyz)
            main = torch.cuda.current_stream(xyz.device)
            fork = torch.cuda.Event()
            fork.record(main)
        joins = []
        for k, (grouper, mlp, idx) in enumerate(zip(self.groupers, self.mlps, indices)):
            if side is not None and k > 0:
                st = side[k - 1]
                st.wait_event(fork)
                with torch.cuda.stream(st):
                    self._scale_inference(k, grouper, xyz, new_xyz, features, idx, out=cat_out, out_coff=offs[k])
                    done = torch.cuda.Event()
                    done.record(st)
                for t in (xyz, new_xyz, features, idx, cat_out):
                    if t is not None:
                        t.record_stream(st)
                joins.append(done)
                continue
            if all_layers:
                # inference: every layer is one tensor-core launch; the last one also max-pools over nsample
                self._scale_inference(k, grouper, xyz, new_xyz, features, idx, out=cat_out, out_coff=offs[k])
                continue
            grouped = grouper(xyz, new_xyz, features, idx=idx)  # (B, 3+C, npoint, nsample)
            B, C, M, K = grouped.shape
            if fused and fused_mlp.supported(M * K, K):   # GroupAll
                pooled.append(self._folded(k)(grouped.view(B, C, M * K), pool=K))
                continue
            if self.pool_method == 'max_pool' and grouped.dim() == 4 and train_mlp.enabled_for(mlp, grouped, K):
                # training: every layer (GEMM + batch-statistics BatchNorm + ReLU, the last one with the max-pool) and its
                # backward on this library's kernels, on the channel-major tensor the grouping kernel has just written
                pooled.append(train_mlp.shared_mlp_train(mlp, grouped.view(B, C, M * K), pool=K))
                continue
            if _train_channels_last(grouped):
                # training on the PyTorch path (WS3D_TRAIN_MLP=0): NHWC activations -- the 1x1 convolutions become plain GEMMs and BatchNorm reduces over the
                # contiguous channel vectors instead of cuDNN's one-thread-block-per-channel NCHW kernels
                grouped = grouped.contiguous(memory_format=torch.channels_last)
            grouped = mlp(grouped)
            if self.pool_method == 'max_pool':
                grouped = F.max_pool2d(grouped, kernel_size=[1, grouped.size(3)])
            elif self.pool_method == 'avg_pool':
                grouped = F.avg_pool2d(grouped, kernel_size=[1, grouped.size(3)])
            else:
                raise NotImplementedError
            pooled.append(grouped.squeeze(-1).contiguous())
        if all_layers:
            for ev in joins:
                main.wait_event(ev)
            return (new_xyz, cat_out, None) if want_rows else (new_xyz, cat_out)
        res = torch.cat(pooled, dim=1)
        return (new_xyz, res, None) if want_rows else (new_xyz, res)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Set abstraction with multi-scale grouping (reference :58-92)."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]], bn: bool = True,
                 use_xyz: bool = True, pool_method='max_pool', instance_norm=False):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz)
                                 if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            if use_xyz:
                spec[0] += 3  # in place, like the reference (:86-87): callers observe the widened spec
            self.mlps.append(pt_utils.SharedMLP(spec, bn=bn, instance_norm=instance_norm))
        self.pool_method = pool_method


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction (reference :95-113)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pool_method='max_pool', instance_norm=False):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz,
                         pool_method=pool_method, instance_norm=instance_norm)


def sa_stack_forward(sa_modules, xyz: torch.Tensor, features: Optional[torch.Tensor]):
    """A chain of set-abstraction levels (the four levels of the Stage-2 network, lib/net/rcnn_net.py:40-58 / its forward loop)
    -> (xyz, features) of the last one.  Same results as calling the modules one after the other; in inference a level that
    runs as the fused kernel also hands its output to the next fused level as point-major operand rows, so only the first
    level packs rows from the channel-major input."""
    mods = list(sa_modules)
    rows = None
    for k, sa in enumerate(mods):
        c_out = sum(m[-1].conv.out_channels for m in sa.mlps)
        c_in = 0 if features is None else features.shape[1]
        chain = (fused_mlp.enabled_for(sa) and xyz.is_cuda and sa.fused_scales_eligible(c_in) and k + 1 < len(mods)
                 and mods[k + 1].fused_scales_eligible(c_out))
        if chain:
            if rows is None and c_in >= 4:
                rows = pointnet2_utils.pack_rows(xyz, features)
            if rows is not None:
                xyz, features, rows = sa(xyz, features, rows=rows, want_rows=True)
                continue
        xyz, features = sa(xyz, features, rows=rows)
        rows = None
    return xyz, features


class PointnetFPModule(nn.Module):
    """Feature propagation (reference :116-156): three_nn -> inverse-distance weights ->
    three_interpolate -> concat skip features -> SharedMLP."""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)

    @staticmethod
    def interpolation_weights(unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """three_nn + inverse-distance weights (reference :139-144).  Depends on coordinates only, so a caller may
        compute it ahead of time (models.Pointnet2MSG does, on a side stream) and pass it as `nn=`."""
        if unknown.is_cuda and unknown.is_contiguous() and known.is_contiguous():
            return pointnet2_utils.three_nn_weights(unknown, known)   # one launch instead of six
        dist, idx = pointnet2_utils.three_nn(unknown, known)
        dist_recip = 1.0 / (dist + 1e-8)
        norm = torch.sum(dist_recip, dim=2, keepdim=True)
        return idx, dist_recip / norm

    def forward(self, unknown: torch.Tensor, known: Optional[torch.Tensor], unknow_feats: Optional[torch.Tensor],
                known_feats: torch.Tensor, nn: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
        if known is not None:
            idx, weight = nn if nn is not None else self.interpolation_weights(unknown, known)
            c_skip = 0 if unknow_feats is None else unknow_feats.shape[1]
            if (fused_mlp.enabled_for(self) and known_feats.is_cuda and known_feats.is_contiguous() and c_skip <= 1
                    and (unknow_feats is None or unknow_feats.is_contiguous())
                    and os.environ.get("WS3D_FP_PREMUL", "1") != "0"
                    and fused_mlp.FoldedFPFirstLayer.eligible(self.mlp[0], known_feats.shape[1], c_skip, known_feats.shape[2],
                                                              unknown.shape[1])):
                # inference, thin skip input: first convolution on the known points, its product interpolated, the skip
                # term + shift + ReLU in the interpolation epilogue; the remaining layers as usual
                cache = self.__dict__.setdefault("_premul", {})
                key = (known_feats.shape[1], c_skip)
                if key not in cache:
                    rest = list(self.mlp)[1:]
                    cache[key] = (fused_mlp.FoldedFPFirstLayer(self.mlp[0], *key),
                                  fused_mlp.FoldedMLP(torch.nn.Sequential(*rest)) if rest else None)
                first, rest = cache[key]
                x = first(known_feats, unknow_feats, idx, weight, unknown.shape[1], round_out=rest is not None)
                return rest(x) if rest is not None else x
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        if fused_mlp.enabled_for(self) and fused_mlp.supported(unknown.size(1), 0) and interpolated.is_contiguous():
            # inference: the skip features are read in place as the second K range (no torch.cat)
            split = (interpolated.shape[1], 0 if unknow_feats is None else unknow_feats.shape[1])
            folded = self.__dict__.get("_folded_mlp")
            if folded is None or folded._first_split != split:
                folded = self.__dict__["_folded_mlp"] = fused_mlp.FoldedMLP(self.mlp, first_split=split)
            skip = None if unknow_feats is None else unknow_feats.contiguous()
            return folded(interpolated, skip)
        if interpolated.dim() == 3 and train_mlp.enabled_for(self.mlp, interpolated):
            # training: the skip features are the second input of the first layer (no torch.cat), see train_mlp
            return train_mlp.shared_mlp_train(self.mlp, interpolated.contiguous(), unknow_feats)
        new_features = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        new_features = new_features.unsqueeze(-1)
        if _train_channels_last(new_features):
            new_features = new_features.contiguous(memory_format=torch.channels_last)
        return self.mlp(new_features).squeeze(-1).contiguous()
