"""The BASELINE.json workloads besides the backbone forward, as callable pieces (bench.py and tools/ share them):

  config 1   one set-abstraction layer's irregular ops on ONE cloud (FPS 16384 -> 4096, ball query r = 0.8, K = 32)
  config 3   the Stage-1 RPN training step (forward, labels, loss, backward, gradient all-reduce over NCCL when world > 1, Adam)
  config 4   rotated BEV IoU + NMS on 16384 boxes and roipool3d of 16384 boxes x 16384 points
  config 5   the Stage-2 set-abstraction stack on 512 pooled proposals per scene

Everything here runs on this repo's kernels; nothing imports oracle/.
"""
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import label_utils, models, native, pointnet2_utils, synth, train_functions
from .pointnet2_modules import PointnetSAModule


def _median_ms(fn: Callable, iters: int, flush: Optional[torch.Tensor], warmup: int = 3) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(0)      # L2 flush between iterations, outside the event pair
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


# ---- config 1 -------------------------------------------------------------------------------------
def single_sa_layer(dev, flush=None, iters: int = 20, npoint: int = 4096, radius: float = 0.8, nsample: int = 32) -> Dict:
    """FPS + gather of the samples + ball query (+ the full QueryAndGroup variant) on one synthetic 16384 x 4 cloud."""
    pts = torch.from_numpy(synth.make_batch(1, 16384)).to(dev)
    xyz = pts[..., :3].contiguous()
    feat = pts[..., 3:].transpose(1, 2).contiguous()
    out = {}

    def ops():
        idx, new_xyz = pointnet2_utils.sample_and_gather(xyz, npoint)
        return idx, new_xyz, pointnet2_utils.ball_query(radius, nsample, xyz, new_xyz)

    def full():
        idx, new_xyz, bq = ops()
        return pointnet2_utils.group_concat(xyz, new_xyz, feat, bq, True)

    with torch.no_grad():
        ms_ops = _median_ms(ops, iters, flush)
        ms_full = _median_ms(full, iters, flush)
        idx, new_xyz, bq = ops()
    n = xyz.shape[1]
    out["ms"] = ms_ops
    out["Mpoints_per_s"] = n / ms_ops / 1e3
    out["full_grouping_ms"] = ms_full
    out["full_grouping_Mpoints_per_s"] = n / ms_full / 1e3
    out["alg_bytes"] = (12 * n + 4 * npoint) + (12 * n + 12 * npoint + 4 * npoint * nsample)
    out["tensors"] = (pts, idx, bq)
    return out


# ---- config 3 -------------------------------------------------------------------------------------
def _plan_tensors(plan: Dict):
    out = list(plan["xyz"])
    for scales in plan["idx"]:
        out += [t for t in scales if t is not None]
    for idx, w in plan["nn"]:
        out += [idx, w]
    out.append(plan["xyz0"])
    if plan["feat0"] is not None:
        out.append(plan["feat0"])
    return out


class RpnTrainStep:
    """Stage-1 training on `batch` synthetic scenes per rank and step: RPN forward (training mode: batch statistics),
    Gaussian labels on the GPU (label_utils), get_rpn_loss (train_functions), backward, the data-parallel
    gradient exchange (sharding.FlatGradients: ONE averaged all-reduce over NCCL, the path's only collective), Adam.
    BatchNorm statistics stay per replica, as under the reference's DataParallel / DDP.

    `prefetch=True` (default): sampling, ball queries and interpolation stencils depend on coordinates only
    (Pointnet2MSG.coordinate_phase), so -- as a data loader knows the NEXT batch while the current one trains -- the
    coordinate phase of batch k+1 runs on a side stream in throughput mode (one SM per cloud) beside step k's forward /
    backward, and step k consumes the plan made during step k-1: the 2.7 ms latency chain of the samplers leaves the
    critical path.  Two synthetic batches alternate; a call = TWO steps (A then B), captured once and replayed as one
    CUDA graph (`graph=True`), collective included.  `exchange=False` leaves the collective out (the all-reduce share of
    a step is measured as the difference).  `steps_per_call` tells the caller how many steps a call makes."""

    def __init__(self, batch: int, dev, world: int = 1, rank: int = 0, graph: bool = True, lr: float = 2e-3, exchange: bool = True,
                 prefetch: bool = True):
        from . import sharding
        torch.manual_seed(0)                      # identical initial replicas on every rank
        self.net = models.RPN().to(dev).train()
        self.model = self.net
        self.param_bytes = sum(p.numel() for p in self.net.parameters()) * 4
        self.world, self.batch, self.dev = world, batch, dev
        self.grads = sharding.FlatGradients(self.net.parameters(), world=world)
        self.exchange = bool(exchange) and world > 1
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr, capturable=bool(graph))
        self.prefetch = bool(prefetch)
        self.steps_per_call = 2 if self.prefetch else 1
        self.data = []
        for k in range(self.steps_per_call):
            first = (2 * rank + k) * batch
            pts = torch.from_numpy(synth.make_batch(batch, 16384, first_scene=first)).to(dev)
            gt, cnt = synth.make_gt_boxes(batch, 16384, first_scene=first)
            self.data.append({"pts": pts, "xyz": pts[..., :3].contiguous(), "gt": torch.from_numpy(gt).to(dev),
                              "cnt": torch.from_numpy(cnt).to(dev), "plan": None})
        self.terms = None
        self._graph = None
        self.graphed = False
        self._side = torch.cuda.Stream(device=dev) if self.prefetch else None
        if self.prefetch:
            self.data[0]["plan"] = self._coordinate_phase(self.data[0]["pts"])
            self.data[1]["plan"] = self._coordinate_phase(self.data[1]["pts"])
            torch.cuda.synchronize()
        if graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._eager()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._static_loss = self._eager()
            self.graphed = True

    def _coordinate_phase(self, pts):
        with torch.no_grad():
            prev = native.set_fps_mode(1)          # throughput mode: one SM per cloud, beside the step's wide kernels
            try:
                return self.net.backbone_net.coordinate_phase(pts)
            finally:
                native.set_fps_mode(prev)

    def _one_step(self, d, plan):
        out = self.model({"pts_input": d["pts"]}, plan=plan)
        with torch.no_grad():
            cls_label, reg_label = label_utils.generate_gaussian_training_labels(d["xyz"], d["gt"], d["cnt"])
        loss, self.terms = train_functions.get_rpn_loss(out["rpn_cls"], out["rpn_reg"], cls_label, reg_label)
        self.grads.zero()
        loss.backward()
        if self.exchange:
            self.grads.exchange()
        self.opt.step()
        return loss

    def _eager(self):
        if not self.prefetch:
            return self._one_step(self.data[0], None)
        main = torch.cuda.current_stream(self.dev)
        loss = None
        for cur, nxt in ((0, 1), (1, 0)):
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                fresh = self._coordinate_phase(self.data[nxt]["pts"])          # what the NEXT step will consume
            loss = self._one_step(self.data[cur], self.data[cur]["plan"])
            main.wait_stream(self._side)
            with torch.no_grad():   # hand-over into the buffers the next step (and the next replay) reads
                for dst, src in zip(_plan_tensors(self.data[nxt]["plan"]), _plan_tensors(fresh)):
                    dst.copy_(src)
                    src.record_stream(main)
        return loss

    def __call__(self):
        if self._graph is not None:
            self._graph.replay()
            return self._static_loss
        return self._eager()


# ---- config 4 -------------------------------------------------------------------------------------
def proposals_scene(dev, num_boxes: int = 16384, channels: int = 128):
    scene = synth.make_scene(0)
    boxes3d = synth.make_boxes(scene[:, :3], num_boxes)
    return {"xyz": torch.from_numpy(scene[None, :, :3].copy()).to(dev),
            "boxes3d": torch.from_numpy(boxes3d).to(dev)[None].contiguous(),
            "bev": torch.from_numpy(synth.boxes3d_to_bev(boxes3d)).to(dev),
            "scores": torch.from_numpy(np.random.default_rng(7).random(num_boxes).astype(np.float32)).to(dev),
            "features": torch.randn((1, scene.shape[0], channels), device=dev, generator=torch.Generator(device=dev).manual_seed(3))}


def iou_nms_roipool(dev, flush=None, iters: int = 5, num_boxes: int = 16384, sampled: int = 512) -> Dict:
    """boxes_iou_bev on num_boxes^2 pairs, nms_gpu at the two thresholds of weaklyRPN.yaml, roipool3d at C = 1 and 128."""
    sc = proposals_scene(dev, num_boxes)
    nb, n = num_boxes, sc["xyz"].shape[1]
    res = {"boxes": nb, "points": n}
    bev = sc["bev"]
    ans = torch.empty((nb, nb), device=dev)
    ms = _median_ms(lambda: native.boxes_iou_bev_gpu(bev, bev, ans), iters, flush)
    res["iou_bev"] = {"ms": ms, "Mpairs_per_s": nb * nb / ms / 1e3, "alg_bytes": 40 * nb + 4 * nb * nb,
                      "GBps": (40 * nb + 4 * nb * nb) / ms / 1e6}
    del ans
    sorted_bev = bev[sc["scores"].sort(descending=True)[1]].contiguous()
    for th in (0.85, 0.1):
        keep = torch.zeros(nb, dtype=torch.int64)
        ms_host = _median_ms(lambda: native.nms_gpu(sorted_bev, keep, th), iters, flush)
        ms_dev = _median_ms(lambda: native.nms_device(sorted_bev, th), iters, flush)
        res[f"nms_{th}"] = {"ms_reference_signature": ms_host, "ms_device_keep": ms_dev, "kept": int(native.nms_gpu(sorted_bev, keep, th)),
                            "mask_bytes": 8 * nb * ((nb + 63) // 64)}
    for c in (1, 128):
        f = sc["features"][..., :c].contiguous()
        pooled = torch.zeros((1, nb, sampled, 3 + c), device=dev)
        flag = torch.zeros((1, nb), dtype=torch.int32, device=dev)
        byts = 12 * n + 4 * c * n + 32 * nb + 4 * nb * sampled * (3 + c)
        ms = _median_ms(lambda: native.roipool3d_forward(sc["xyz"], sc["boxes3d"], f, pooled, flag), iters, flush)
        res[f"roipool3d_C{c}"] = {"ms": ms, "alg_bytes": byts, "GBps": byts / ms / 1e6, "empty_boxes": int(flag.sum())}
        del pooled
    res["scene"] = sc
    return res


# ---- config 5 -------------------------------------------------------------------------------------
STAGE2_SA = {"NPOINTS": [256, 128, 32, -1], "RADIUS": [0.2, 0.4, 1.0, 100.0], "NSAMPLE": [16, 32, 64, 64],
             "MLPS": [[128, 128, 128], [128, 128, 128], [128, 128, 256], [256, 256, 512]]}   # weaklyRCNN.yaml:60-77


class Stage2SA(nn.Module):
    """The four set-abstraction levels of the Stage-2 network (lib/net/rcnn_net.py:40-58) at the weaklyRCNN.yaml shapes."""

    def __init__(self, channel_in: int = 128):
        super().__init__()
        self.SA_modules = nn.ModuleList()
        for k in range(len(STAGE2_SA["NPOINTS"])):
            npoint = STAGE2_SA["NPOINTS"][k] if STAGE2_SA["NPOINTS"][k] != -1 else None
            self.SA_modules.append(PointnetSAModule(npoint=npoint, radius=STAGE2_SA["RADIUS"][k], nsample=STAGE2_SA["NSAMPLE"][k],
                                                    mlp=[channel_in] + STAGE2_SA["MLPS"][k], use_xyz=True, bn=True))
            channel_in = STAGE2_SA["MLPS"][k][-1]

    def forward(self, xyz, features):
        from .pointnet2_modules import sa_stack_forward
        return sa_stack_forward(self.SA_modules, xyz, features)[1]


def stage2_stack(dev, rank: int = 0, scenes: int = 1, flush=None, iters: int = 10) -> Dict:
    """512 proposals per scene x 512 points x 128 channels through the Stage-2 SA stack, forward, eval mode."""
    torch.manual_seed(0)
    model = Stage2SA().to(dev).eval()
    B = 512 * scenes
    rng = np.random.default_rng(1234 + rank)
    xyz = torch.from_numpy((rng.normal(0, 1, (B, 512, 3)) * np.array([1.2, 0.6, 2.2])).astype(np.float32)).to(dev)
    feats = torch.randn(B, 128, 512, device=dev)
    with torch.no_grad():
        ms = _median_ms(lambda: model(xyz, feats), iters, flush)
    return {"ms": ms, "proposals_per_gpu": B, "proposals_per_s": B / ms * 1e3}
