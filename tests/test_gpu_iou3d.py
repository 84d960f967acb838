"""GPU parity tests for iou3d: rotated BEV overlap / IoU matrices and NMS keep lists.
vs oracle/_ref (the reference kernels, same libdevice): BIT-EXACT matrices and keep lists.
vs the CPU oracle (host libm trig): <= 1e-5, and identical keep lists away from the threshold."""
import numpy as np
import pytest
import torch

import oracle
from refmods import require_ref

pytestmark = pytest.mark.gpu
dev = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _boxes(rng, n, spread=20.0, mode="random"):
    """(n,5) [x1,y1,x2,y2,ry] car-sized boxes clustered so that many pairs overlap."""
    cx, cy = rng.uniform(-spread, spread, n), rng.uniform(-spread, spread, n)
    l, w = 3.9 + rng.normal(0, 0.3, n), 1.6 + rng.normal(0, 0.1, n)
    ry = rng.uniform(-np.pi, np.pi, n)
    if mode == "axis":        # axis-aligned and shared edges: collinear / degenerate intersections
        ry = rng.choice([0.0, np.pi / 2, np.pi, -np.pi / 2], n)
        cx, cy = np.round(cx), np.round(cy)
        l, w = np.full(n, 4.0), np.full(n, 2.0)
    if mode == "dup":         # exact duplicates and near-duplicates
        half = n // 2
        cx[half:], cy[half:], l[half:], w[half:], ry[half:] = cx[:n - half], cy[:n - half], l[:n - half], w[:n - half], ry[:n - half]
        cx[half:] += rng.choice([0.0, 1e-4, 0.05], n - half)
    return np.stack([cx - l / 2, cy - w / 2, cx + l / 2, cy + w / 2, ry], 1).astype(np.float32)


@pytest.mark.parametrize("na,nb,spread,mode", [(300, 257, 8.0, "random"), (1000, 1000, 25.0, "random"), (128, 128, 4.0, "axis"),
                                               (200, 200, 6.0, "dup"), (1, 1, 1.0, "random"), (65, 3, 2.0, "random")])
def test_overlap_and_iou_matrices(na, nb, spread, mode):
    from ws3d_b200 import iou3d_utils, native
    rng = np.random.default_rng(na * 31 + nb)
    a, b = _boxes(rng, na, spread, mode), _boxes(rng, nb, spread, mode)
    ta, tb = _t(a), _t(b)
    ov = torch.zeros((na, nb), device=dev)
    native.boxes_overlap_bev_gpu(ta, tb, ov)
    iou = iou3d_utils.boxes_iou_bev(ta, tb)
    ref = require_ref("iou3d_cuda")
    if ref is not None:
        rov, riou = torch.zeros_like(ov), torch.zeros_like(iou)
        ref.boxes_overlap_bev_gpu(ta, tb, rov)
        ref.boxes_iou_bev_gpu(ta, tb, riou)
        torch.cuda.synchronize()
        # bit-exact (NaNs, if any, must coincide too)
        assert torch.equal(torch.nan_to_num(ov, nan=-7.0), torch.nan_to_num(rov, nan=-7.0))
        assert torch.equal(torch.nan_to_num(iou, nan=-7.0), torch.nan_to_num(riou, nan=-7.0))
    if mode == "random":  # degenerate configurations are decided by last-ulp trig: GPU oracle only
        np.testing.assert_allclose(ov.cpu().numpy(), oracle.boxes_overlap_bev(a, b), rtol=1e-4, atol=2e-4)
        np.testing.assert_allclose(iou.cpu().numpy(), oracle.boxes_iou_bev(a, b), rtol=1e-4, atol=1e-5)


def test_iou_self_and_symmetry_properties():
    from ws3d_b200 import iou3d_utils
    rng = np.random.default_rng(3)
    a = _boxes(rng, 2048, 30.0)
    ta = _t(a)
    iou = iou3d_utils.boxes_iou_bev(ta, ta)
    d = torch.diagonal(iou)
    assert float((d - 1).abs().max()) < 1e-4                 # IoU(box, box) == 1
    assert float(iou.min()) >= 0 and float(iou.max()) < 1 + 1e-4
    assert float((iou - iou.t()).abs().max()) < 1e-4         # symmetric up to rounding


@pytest.mark.parametrize("n,spread,thresh,mode", [(2000, 10.0, 0.1, "random"), (2000, 10.0, 0.85, "random"), (777, 5.0, 0.5, "random"),
                                                   (64, 3.0, 0.3, "random"), (65, 3.0, 0.3, "random"), (1, 1.0, 0.5, "random"),
                                                   (500, 6.0, 0.7, "dup"), (4096, 12.0, 0.8, "random")])
def test_rotated_nms_keep_list(n, spread, thresh, mode):
    from ws3d_b200 import iou3d_utils, native
    rng = np.random.default_rng(n)
    boxes = _boxes(rng, n, spread, mode)
    scores = rng.permutation(n).astype(np.float32)  # distinct
    tb, ts = _t(boxes), _t(scores)
    keep = iou3d_utils.nms_gpu(tb, ts, thresh)
    order = torch.sort(ts, descending=True)[1]
    sorted_boxes = tb[order].contiguous()
    # reference-signature entry point (CPU int64 keep buffer)
    kbuf = torch.zeros(n, dtype=torch.int64)
    num = native.nms_gpu(sorted_boxes, kbuf, thresh)
    assert torch.equal(order[kbuf[:num].to(dev)], keep)
    ref = require_ref("iou3d_cuda")
    if ref is not None:
        rbuf = torch.zeros(n, dtype=torch.int64)
        rnum = ref.nms_gpu(sorted_boxes, rbuf, thresh)
        assert rnum == num and torch.equal(rbuf[:rnum], kbuf[:num])          # bit-exact keep list
    ok = oracle.nms(sorted_boxes.cpu().numpy(), thresh)
    if ref is None:
        assert abs(len(ok) - num) <= max(2, n // 200)                          # host libm vs libdevice near thresh
    # greedy-NMS invariant: no kept pair overlaps above the threshold
    kept = sorted_boxes[kbuf[:num].to(dev)]
    iou = iou3d_utils.boxes_iou_bev(kept, kept)
    iou.fill_diagonal_(0)
    assert float(iou.max()) <= thresh + 1e-6 if num > 1 else True


@pytest.mark.parametrize("n,thresh", [(3000, 0.5), (130, 0.1), (1, 0.3)])
def test_normal_nms_keep_list(n, thresh):
    from ws3d_b200 import iou3d_utils, native
    rng = np.random.default_rng(n + 1)
    boxes = _boxes(rng, n, 10.0)
    scores = rng.permutation(n).astype(np.float32)
    tb, ts = _t(boxes), _t(scores)
    keep = iou3d_utils.nms_normal_gpu(tb, ts, thresh)
    order = torch.sort(ts, descending=True)[1]
    sorted_boxes = tb[order].contiguous()
    exp = oracle.nms_normal(sorted_boxes.cpu().numpy(), thresh)
    np.testing.assert_array_equal(keep.cpu().numpy(), order.cpu().numpy()[exp])   # no trig: bit-exact vs the CPU oracle
    ref = require_ref("iou3d_cuda")
    if ref is not None:
        rbuf = torch.zeros(n, dtype=torch.int64)
        rnum = ref.nms_normal_gpu(sorted_boxes, rbuf, thresh)
        np.testing.assert_array_equal(rbuf[:rnum].numpy(), exp)


def test_boxes_iou3d_gpu_matches_reference_formula():
    from ws3d_b200 import iou3d_utils, synth
    rng = np.random.default_rng(11)
    pts = synth.make_scene(0)[:, :3]
    a = synth.make_boxes(pts, 300, seed=1)
    b = a[rng.permutation(300)[:200]] + rng.normal(0, 0.2, (200, 7)).astype(np.float32)
    iou2d, iou3d = iou3d_utils.boxes_iou3d_gpu(_t(a), _t(b))
    ov = oracle.boxes_overlap_bev(synth.boxes3d_to_bev(a), synth.boxes3d_to_bev(b))
    hmin = np.maximum((a[:, 1] - a[:, 3])[:, None], (b[:, 1] - b[:, 3])[None])
    hmax = np.minimum(a[:, 1][:, None], b[:, 1][None])
    oh = np.clip(hmax - hmin, 0, None)
    va, vb = (a[:, 3] * a[:, 4] * a[:, 5])[:, None], (b[:, 3] * b[:, 4] * b[:, 5])[None]
    exp3d = ov * oh / np.clip(va + vb - ov * oh, 1e-7, None)
    np.testing.assert_allclose(iou3d.cpu().numpy(), exp3d, rtol=1e-4, atol=1e-5)
    assert iou2d.shape == (300, 200)


def test_nms_config4_size_equals_reference_kernels():
    """BASELINE configs[3] size: 16384 rotated boxes, both thresholds of weaklyRPN.yaml -- keep lists equal to the
    reference's nms_gpu / nms_normal_gpu, and the all-device variant equal to the reference-signature one."""
    from ws3d_b200 import native, synth
    scene = synth.make_scene(0)
    boxes3d = synth.make_boxes(scene[:, :3], 16384)
    bev = torch.from_numpy(synth.boxes3d_to_bev(boxes3d)).to(dev)
    scores = torch.from_numpy(np.random.default_rng(7).random(16384).astype(np.float32)).to(dev)
    sb = bev[scores.sort(descending=True)[1]].contiguous()
    ref = require_ref("iou3d_cuda")
    for th in (0.85, 0.1):
        for mine, theirs in ((native.nms_gpu, ref.nms_gpu), (native.nms_normal_gpu, ref.nms_normal_gpu)):
            kb, rb = torch.zeros(16384, dtype=torch.int64), torch.zeros(16384, dtype=torch.int64)
            num, rnum = mine(sb, kb, th), theirs(sb, rb, th)
            assert num == rnum and torch.equal(kb[:num], rb[:rnum]), (th, mine.__name__)
        keep, cnt = native.nms_device(sb, th)
        kb = torch.zeros(16384, dtype=torch.int64)
        num = native.nms_gpu(sb, kb, th)
        assert int(cnt.item()) == num and torch.equal(keep[:num].cpu(), kb[:num])
