"""The callers of the hot path, at the reference's shapes: the Stage-1 PointNet++-MSG backbone
(lib/net/pointnet2_msg.py) and the RPN heads on top of it (lib/net/rpn.py:20-45,67-81).

Host side stays PyTorch (north star): these are ordinary nn.Modules whose SA / FP layers call the
B200 ops.  Structure and parameter names follow the reference so its checkpoints load unchanged
(`backbone_net.SA_modules.0.mlps.0.layer0.conv.weight`, ...).  The network shape comes from
tools/cfgs/weaklyRPN.yaml:43-56 (= lib/config.py:57-70), restated here as RPN_SA_CONFIG.
"""
import copy
import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import fused_mlp
from . import pointnet2_utils
from . import train_mlp
from . import pytorch_utils as pt_utils
from .pointnet2_modules import PointnetFPModule, PointnetSAModuleMSG

RPN_SA_CONFIG = {
    "NPOINTS": [4096, 1024, 256, 64],
    "RADIUS": [[0.1, 0.5], [0.5, 1.0], [1.0, 2.0], [2.0, 4.0]],
    "NSAMPLE": [[16, 32], [16, 32], [16, 32], [16, 32]],
    "MLPS": [[[16, 16, 32], [32, 32, 64]],
             [[64, 64, 128], [64, 96, 128]],
             [[128, 196, 256], [128, 196, 256]],
             [[256, 256, 512], [256, 384, 512]]],
}
RPN_FP_MLPS = [[128, 128], [256, 256], [512, 512], [512, 512]]
RPN_CLS_FC = [128]
RPN_REG_FC = [128]
RPN_DP_RATIO = 0.5
RPN_LOC_SCOPE = 4.0      # weaklyRPN.yaml:37
RPN_LOC_BIN_SIZE = 0.8   # weaklyRPN.yaml:38
RPN_NUM_POINTS = 16384


class Pointnet2MSG(nn.Module):
    """4 x PointnetSAModuleMSG + 4 x PointnetFPModule (lib/net/pointnet2_msg.py:11-70)."""

    def __init__(self, input_channels: int = 1, use_xyz: bool = True, sa_config=None, fp_mlps=None, bn: bool = True):
        super().__init__()
        sa = copy.deepcopy(sa_config or RPN_SA_CONFIG)
        fp = copy.deepcopy(fp_mlps or RPN_FP_MLPS)
        self.SA_modules = nn.ModuleList()
        channel_in = input_channels
        skip_channels = [input_channels]
        channel_out = channel_in
        for k in range(len(sa["NPOINTS"])):
            mlps = [[channel_in] + list(m) for m in sa["MLPS"][k]]
            channel_out = sum(m[-1] for m in mlps)
            self.SA_modules.append(PointnetSAModuleMSG(npoint=sa["NPOINTS"][k], radii=sa["RADIUS"][k],
                                                       nsamples=sa["NSAMPLE"][k], mlps=mlps, use_xyz=use_xyz, bn=bn))
            skip_channels.append(channel_out)
            channel_in = channel_out
        self.FP_modules = nn.ModuleList()
        for k in range(len(fp)):
            pre = fp[k + 1][-1] if k + 1 < len(fp) else channel_out
            self.FP_modules.append(PointnetFPModule(mlp=[pre + skip_channels[k]] + fp[k], bn=True))

    @staticmethod
    def _break_up_pc(pc):
        """(B,N,3+C) -> xyz (B,N,3), features (B,C,N) or None (lib/net/pointnet2_msg.py:52-60); one launch on the GPU."""
        if pc.is_cuda and pc.is_contiguous() and pc.dtype == torch.float32 and not (torch.is_grad_enabled() and pc.requires_grad):
            from . import native
            B, N, C = pc.shape[0], pc.shape[1], pc.shape[2] - 3
            xyz = torch.empty((B, N, 3), dtype=torch.float32, device=pc.device)
            features = torch.empty((B, C, N), dtype=torch.float32, device=pc.device) if C > 0 else None
            native.split_pointcloud(pc, xyz, features)
            return xyz, features
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def _sa_level(self, k: int, pointcloud: torch.Tensor, rows, xyz, feats, new_xyz=None, indices=None):
        """Set-abstraction level k -> (new_xyz, new_features, rows_out).  `rows` are the point-major operand rows
        [xyz | features | zeros] the fused kernels gather from in inference: level 0 reads the (B,N,3+C) input cloud itself
        when its rows are 16-byte multiples, and a level emits its output in that layout (rows_out) when the next level
        runs fused too (pointnet2_modules.fused_scales_eligible); None wherever that does not apply."""
        sa = self.SA_modules[k]
        want = False
        if fused_mlp.enabled_for(sa) and pointcloud.is_cuda:
            if k == 0:
                ok = (pointcloud.is_contiguous() and pointcloud.dtype == torch.float32 and pointcloud.shape[-1] % 4 == 0
                      and sa.fused_scales_eligible(pointcloud.shape[-1] - 3))
                rows = pointcloud if ok else None
            if rows is not None and k + 1 < len(self.SA_modules):
                want = self.SA_modules[k + 1].fused_scales_eligible(sum(m[-1].conv.out_channels for m in sa.mlps))
        else:
            rows = None
        if rows is None:
            nx, nf = sa(xyz, feats, new_xyz=new_xyz, indices=indices)
            return nx, nf, None
        if want:
            return sa(xyz, feats, new_xyz=new_xyz, indices=indices, rows=rows, want_rows=True)
        nx, nf = sa(xyz, feats, new_xyz=new_xyz, indices=indices, rows=rows)
        return nx, nf, None

    def sample_first_level(self, pointcloud: torch.Tensor) -> torch.Tensor:
        """Level-1 sampling alone: (B,N,3+C) -> new_xyz (B, npoint_1, 3).  It depends on coordinates only and is the
        head of every dependency chain of the forward pass, so a caller that knows the NEXT batch can compute it
        while the current batch is still in its grouping / MLP layers (graphs.PipelinedBackboneRunner) and hand it
        to forward(..., first_samples=)."""
        xyz = pointcloud[..., 0:3].contiguous()
        return pointnet2_utils.sample_and_gather(xyz, self.SA_modules[0].npoint)[1]

    def coordinate_phase(self, pointcloud: torch.Tensor) -> dict:
        """Everything of the forward pass that depends on coordinates only: the four FPS levels, the ball queries of
        every scale and the three-nearest-neighbour interpolation stencils.  It is a chain of latency-bound kernels
        that needs few SMs and no bandwidth, so a caller with a stream of batches runs it AHEAD of the feature phase,
        beside the previous batches' feature phases (graphs.StreamedBackboneRunner)."""
        xyz, features = self._break_up_pc(pointcloud)
        if not xyz.is_cuda or os.environ.get("WS3D_COORD_SIDE", "1") == "0":
            l_xyz, idx = [xyz], []
            for sa in self.SA_modules:
                _, nx = pointnet2_utils.sample_and_gather(l_xyz[-1], sa.npoint)
                idx.append(sa._neighbour_indices(l_xyz[-1], nx))
                l_xyz.append(nx)
            nn_ = [PointnetFPModule.interpolation_weights(l_xyz[i], l_xyz[i + 1]) for i in range(len(self.FP_modules))]
            return {"xyz": l_xyz[1:], "idx": idx, "nn": nn_, "xyz0": xyz, "feat0": features}
        # The sampling levels are one dependency chain (level k + 1 samples level k's samples); the ball queries and the
        # interpolation stencils only hang off it.  They run on a side stream as soon as their level exists, so the phase is as long
        # as its four FPS launches (what a batch that fills an empty pipeline waits for; in a full pipeline the order is irrelevant).
        dev = xyz.device
        main = torch.cuda.current_stream(dev)
        side = self.__dict__.get("_coord_side_stream")
        if side is None or side.device != dev:
            side = self.__dict__["_coord_side_stream"] = torch.cuda.Stream(device=dev)
        n_fp = len(self.FP_modules)
        l_xyz, idx, nn_ = [xyz], [], [None] * n_fp
        xyz.record_stream(side)
        for k, sa in enumerate(self.SA_modules):
            _, nx = pointnet2_utils.sample_and_gather(l_xyz[-1], sa.npoint)
            nx.record_stream(side)
            level_done = torch.cuda.Event()
            level_done.record(main)
            with torch.cuda.stream(side):
                side.wait_event(level_done)
                ind = sa._neighbour_indices(l_xyz[-1], nx)
                for t in ind:
                    t.record_stream(main)
                idx.append(ind)
                if k < n_fp:
                    nn_[k] = PointnetFPModule.interpolation_weights(l_xyz[-1], nx)
                    for t in nn_[k]:
                        t.record_stream(main)
            l_xyz.append(nx)
        main.wait_stream(side)
        return {"xyz": l_xyz[1:], "idx": idx, "nn": nn_, "xyz0": xyz, "feat0": features}

    def feature_phase(self, pointcloud: torch.Tensor, plan: dict):
        """The rest: grouping + MLPs of the SA levels, interpolation + MLPs of the FP levels, on precomputed samples,
        neighbour indices and stencils.  Same kernels and arithmetic as forward()."""
        if "xyz0" in plan:   # the coordinate phase has already split the cloud
            xyz, features = plan["xyz0"], plan["feat0"]
        else:
            xyz, features = self._break_up_pc(pointcloud)
        l_xyz, l_features = [xyz] + list(plan["xyz"]), [features]
        rows = None
        for k in range(len(self.SA_modules)):
            _, nf, rows = self._sa_level(k, pointcloud, rows, l_xyz[k], l_features[k], new_xyz=l_xyz[k + 1], indices=plan["idx"][k])
            l_features.append(nf)
        for i in range(len(self.FP_modules) - 1, -1, -1):
            l_features[i] = self.FP_modules[i](l_xyz[i], l_xyz[i + 1], l_features[i], l_features[i + 1], nn=plan["nn"][i])
        return l_xyz[0], l_features[0]

    def _forward_two_streams(self, pointcloud, xyz, features, first_samples=None):
        """Same computation as forward(), scheduled on two CUDA streams.

        Sampling (FPS) and the interpolation stencils (three_nn + weights) depend on coordinates only, and FPS is a
        latency chain that occupies few SMs below the first level.  They run ahead on a side stream while the main
        stream does ball query / grouping / MLPs of the levels whose samples are already known; events order the
        two, `record_stream` keeps the caching allocator from recycling a tensor the other stream still reads.
        """
        main = torch.cuda.current_stream(xyz.device)
        side = self.__dict__.get("_side_stream")
        if side is None or side.device != xyz.device:
            # high priority: the sampling chain is latency-bound and narrow; it must not queue behind wide kernels
            side = self.__dict__["_side_stream"] = torch.cuda.Stream(device=xyz.device, priority=-1)
        start = torch.cuda.Event()
        start.record(main)
        l_xyz, fps_done, stencil, stencil_done = [xyz], [], {}, {}
        with torch.cuda.stream(side):
            side.wait_event(start)
            for k, sa in enumerate(self.SA_modules):
                if k == 0 and first_samples is not None:
                    nx = first_samples
                else:
                    _, nx = pointnet2_utils.sample_and_gather(l_xyz[-1], sa.npoint)
                nx.record_stream(main)
                l_xyz.append(nx)
                ev = torch.cuda.Event()
                ev.record(side)
                fps_done.append(ev)
            for i in range(len(self.FP_modules) - 1, -1, -1):  # deepest level is needed first
                idx, weight = PointnetFPModule.interpolation_weights(l_xyz[i], l_xyz[i + 1])
                idx.record_stream(main)
                weight.record_stream(main)
                stencil[i] = (idx, weight)
                ev = torch.cuda.Event()
                ev.record(side)
                stencil_done[i] = ev
        l_features, rows = [features], None
        for i in range(len(self.SA_modules)):
            main.wait_event(fps_done[i])
            _, nf, rows = self._sa_level(i, pointcloud, rows, l_xyz[i], l_features[i], new_xyz=l_xyz[i + 1])
            l_features.append(nf)
        for i in range(len(self.FP_modules) - 1, -1, -1):
            main.wait_event(stencil_done[i])
            l_features[i] = self.FP_modules[i](l_xyz[i], l_xyz[i + 1], l_features[i], l_features[i + 1], nn=stencil[i])
        xyz.record_stream(side)
        return l_xyz[0], l_features[0]

    def forward(self, pointcloud: torch.Tensor, first_samples: Optional[torch.Tensor] = None):
        """pointcloud (B,N,3+C) -> (xyz (B,N,3), per-point features (B,128,N)).
        `first_samples`: level-1 FPS output for this batch if the caller already has it (sample_first_level)."""
        xyz, features = self._break_up_pc(pointcloud)
        if (xyz.is_cuda and os.environ.get("WS3D_TWO_STREAMS", "1") != "0" and len(self.FP_modules) == len(self.SA_modules)
                and all(sa.npoint is not None for sa in self.SA_modules)):
            return self._forward_two_streams(pointcloud, xyz, features, first_samples)
        l_xyz, l_features, rows = [xyz], [features], None
        for k in range(len(self.SA_modules)):
            nx, nf, rows = self._sa_level(k, pointcloud, rows, l_xyz[-1], l_features[-1], new_xyz=first_samples if k == 0 else None)
            l_xyz.append(nx)
            l_features.append(nf)
        for i in range(-1, -(len(self.FP_modules) + 1), -1):
            l_features[i - 1] = self.FP_modules[i](l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i])
        return l_xyz[0], l_features[0]


class RPN(nn.Module):
    """Backbone + per-point classification / bin-regression heads (lib/net/rpn.py:10-81)."""

    def __init__(self, use_xyz: bool = True, use_intensity: bool = True, focal_init: bool = True):
        super().__init__()
        self.backbone_net = Pointnet2MSG(input_channels=int(use_intensity), use_xyz=use_xyz)

        def head(fc, out_channels):
            layers, pre = [], RPN_FP_MLPS[0][-1]
            for width in fc:
                layers.append(pt_utils.Conv1d(pre, width, bn=True))
                pre = width
            layers.append(pt_utils.Conv1d(pre, out_channels, activation=None))
            if RPN_DP_RATIO >= 0:
                layers.insert(1, nn.Dropout(RPN_DP_RATIO))
            return nn.Sequential(*layers)

        per_loc_bin_num = int(RPN_LOC_SCOPE / RPN_LOC_BIN_SIZE) * 2
        self.reg_channel = per_loc_bin_num * 4
        self.rpn_cls_layer = head(RPN_CLS_FC, 1)
        self.rpn_reg_layer = head(RPN_REG_FC, self.reg_channel)
        if focal_init:  # rpn.py:60-65
            pi = 0.01
            nn.init.constant_(self.rpn_cls_layer[2].conv.bias, -np.log((1 - pi) / pi))
        nn.init.normal_(self.rpn_reg_layer[-1].conv.weight, mean=0, std=0.001)

    def _head(self, name: str, x: torch.Tensor) -> torch.Tensor:
        """A per-point head (Conv1d + BN + ReLU, Dropout, Conv1d).  Inference: the tensor-core layer kernel
        (dropout is the identity in eval mode); otherwise the PyTorch modules."""
        seq = getattr(self, name)
        if fused_mlp.enabled_for(self) and x.is_cuda and fused_mlp.supported(x.shape[2], 0) and x.is_contiguous():
            cache = self.__dict__.setdefault("_folded_heads", {})
            if name not in cache:
                cache[name] = fused_mlp.FoldedMLP(nn.Sequential(*[m for m in seq if not isinstance(m, nn.Dropout)]))
            return cache[name](x)
        if x.dim() == 3 and train_mlp.enabled_for(seq, x):
            return train_mlp.shared_mlp_train(seq, x.contiguous())     # training: this library's layer kernels + torch's dropout
        return seq(x)

    def forward(self, input_data, first_samples: Optional[torch.Tensor] = None, plan: Optional[dict] = None):
        pts_input = input_data['pts_input'] if isinstance(input_data, dict) else input_data
        if plan is not None:
            backbone_xyz, backbone_features = self.backbone_net.feature_phase(pts_input, plan)
        else:
            backbone_xyz, backbone_features = self.backbone_net(pts_input, first_samples=first_samples)
        if (backbone_features.is_cuda and fused_mlp.enabled_for(self) and backbone_features.is_contiguous()
                and fused_mlp.supported(backbone_features.shape[2], 0) and len(self.rpn_cls_layer) == len(self.rpn_reg_layer)):
            # inference: both heads as one chain of two launches (stacked first layers, block-diagonal second layers)
            heads = self.__dict__.get("_folded_heads2")
            if heads is None:
                heads = self.__dict__["_folded_heads2"] = fused_mlp.FoldedHeads([self.rpn_cls_layer, self.rpn_reg_layer])
            both = heads(backbone_features)                                   # (B, 1 + reg_channel, N)
            n_cls = heads.out_channels[0]
            rpn_cls = both[:, :n_cls].transpose(1, 2).contiguous()            # (B, N, 1)
            rpn_reg = both[:, n_cls:].transpose(1, 2).contiguous()            # (B, N, reg_channel)
        else:
            rpn_cls = self._head("rpn_cls_layer", backbone_features).transpose(1, 2).contiguous()
            rpn_reg = self._head("rpn_reg_layer", backbone_features).transpose(1, 2).contiguous()
        return {'rpn_cls': rpn_cls, 'rpn_reg': rpn_reg, 'backbone_xyz': backbone_xyz,
                'backbone_features': backbone_features}
