"""Boundary proof: the reference's own Python -- pointnet2_utils.py, pointnet2_modules.py,
iou3d_utils.py, roipool3d_utils.py, UNMODIFIED (oracle/_ref/refpy.zip) -- run twice on the same inputs and weights: over
this repo's drop-in modules (ws3d_b200.dropin.*, i.e. libws3d_ops.so through the C ABI) and over the reference's
extensions recompiled for sm_100a (oracle/_ref/*.so).  The two must agree: bit-exact indices / keep lists / pooled rows,
and identical features (the MLPs are the same PyTorch modules with TF32 off)."""
import importlib

import numpy as np
import pytest
import torch

from refmods import REFPY_ZIP, load_reference_python, require_ref, require_reference_python
from ws3d_b200 import synth


def _dropins():
    return {name: importlib.import_module(f"ws3d_b200.dropin.{name}") for name in ("pointnet2_cuda", "iou3d_cuda", "roipool3d_cuda")}


def test_reference_python_imports_over_the_dropins_without_a_gpu():
    """CPU: the unmodified wrappers import against the drop-in function tables (every name they bind exists)."""
    import os
    if not os.path.exists(REFPY_ZIP):
        pytest.skip("oracle/_ref/refpy.zip not staged (no /root/reference on this machine)")
    ns = load_reference_python(_dropins())
    assert ns.pointnet2_utils.pointnet2.__name__.startswith("ws3d_b200.dropin")
    assert ns.iou3d_utils.iou3d_cuda.__name__.startswith("ws3d_b200.dropin")
    assert ns.roipool3d_utils.roipool3d_cuda.__name__.startswith("ws3d_b200.dropin")
    for fn in ("furthest_point_sample", "gather_operation", "three_nn", "three_interpolate", "grouping_operation", "ball_query"):
        assert callable(getattr(ns.pointnet2_utils, fn))
    import sys
    assert "pointnet2_lib" not in sys.modules and "lib" not in sys.modules    # nothing leaks into the interpreter


@pytest.fixture(scope="module")
def both():
    assert torch.cuda.is_available()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    mine = require_reference_python(_dropins())
    ref = require_reference_python({name: require_ref(name) for name in ("pointnet2_cuda", "iou3d_cuda", "roipool3d_cuda")})
    yield mine, ref
    torch.backends.cudnn.allow_tf32 = True


@pytest.mark.gpu
def test_reference_sa_and_fp_modules_over_dropins_equal_reference_kernels(both):
    mine, ref = both
    dev = "cuda:0"
    pts = torch.from_numpy(synth.make_batch(2, 4096)).to(dev)
    xyz = pts[..., :3].contiguous()
    feats = pts[..., 3:].transpose(1, 2).contiguous()
    outs = []
    for ns in (mine, ref):
        torch.manual_seed(5)
        sa = ns.pointnet2_modules.PointnetSAModuleMSG(npoint=1024, radii=[0.5, 1.0], nsamples=[16, 32],
                                                      mlps=[[1, 16, 16, 32], [1, 32, 32, 64]], use_xyz=True, bn=True).to(dev).eval()
        fp = ns.pointnet2_modules.PointnetFPModule(mlp=[96 + 1, 64, 64]).to(dev).eval()
        with torch.no_grad():
            new_xyz, new_feats = sa(xyz, feats)
            up = fp(xyz, new_xyz, feats, new_feats)
        # and the gradient ops through the reference's autograd Functions
        f = feats.clone().requires_grad_(True)
        idx = ns.pointnet2_utils.ball_query(1.0, 32, xyz, new_xyz)
        g = ns.pointnet2_utils.grouping_operation(f, idx)
        dist, nn_idx = ns.pointnet2_utils.three_nn(xyz, new_xyz)
        w = 1.0 / (dist + 1e-8)
        w = w / w.sum(dim=2, keepdim=True)
        nf = new_feats.clone().requires_grad_(True)
        interp = ns.pointnet2_utils.three_interpolate(nf, nn_idx, w)
        (g.sum() * 0.5 + (interp * interp).sum()).backward()
        outs.append((new_xyz, new_feats, up, idx, nn_idx, dist, f.grad.clone(), nf.grad.clone()))
    names = ("new_xyz", "sa features", "fp features", "ball_query idx", "three_nn idx", "three_nn dist")
    for name, a, b in zip(names, outs[0], outs[1]):
        assert torch.equal(a, b), name
    torch.testing.assert_close(outs[0][6], outs[1][6], rtol=1e-5, atol=1e-5)     # atomics: summation order
    torch.testing.assert_close(outs[0][7], outs[1][7], rtol=1e-5, atol=1e-4)


@pytest.mark.gpu
def test_reference_iou3d_and_roipool_wrappers_over_dropins_equal_reference_kernels(both):
    mine, ref = both
    dev = "cuda:0"
    scene = synth.make_scene(3)
    boxes3d = torch.from_numpy(synth.make_boxes(scene[:, :3], 1500)).to(dev)
    scores = torch.from_numpy(np.random.default_rng(1).random(1500).astype(np.float32)).to(dev)
    pts = torch.from_numpy(scene[None, :, :3].copy()).to(dev)
    feat = torch.from_numpy(scene[None, :, 3:].copy()).to(dev)
    res = []
    for ns in (mine, ref):
        bev = ns.iou3d_utils.kitti_utils.boxes3d_to_bev_torch(boxes3d)
        keep = ns.iou3d_utils.nms_gpu(bev, scores, 0.85)
        keep_n = ns.iou3d_utils.nms_normal_gpu(bev, scores, 0.8)
        iou_bev = ns.iou3d_utils.boxes_iou_bev(bev[:300].contiguous(), bev[300:700].contiguous())
        iou2d, iou3d = ns.iou3d_utils.boxes_iou3d_gpu(boxes3d[:200].contiguous(), boxes3d[200:500].contiguous())
        pooled, flag = ns.roipool3d_utils.roipool3d_gpu(pts, feat, boxes3d[None, :256].contiguous(), 1.0, sampled_pt_num=512)
        res.append((keep, keep_n, iou_bev, iou2d, iou3d, pooled, flag))
    for name, a, b in zip(("nms_gpu keep", "nms_normal_gpu keep", "boxes_iou_bev", "iou2d", "iou3d", "pooled", "empty flag"), res[0], res[1]):
        assert torch.equal(a, b), name
