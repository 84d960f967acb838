"""GPU parity for the SURVEY.md section-8 "next" rows f2 / f3, through the C ABI:
aligned-pair IoU3D (diagonal of boxes_iou3d_gpu), radius NMS and the cylinder crop of tools/eval_auto.py.
Checked against the CPU oracle, the fixtures made by executing the reference's Python (tests/golden/next_rows.npz),
the full-matrix path of this library, and size-independent properties at 16384 candidates."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
dev = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "next_rows.npz")


def _dist(a, b):
    """lib/utils/distance.py:3 in torch float32 (one kernel per operation, as the reference runs it)."""
    return torch.sqrt(torch.sum((a[None, :] - b[:, None]) ** 2, dim=2))


def test_iou3d_aligned_matches_golden_oracle_and_full_matrix_diagonal():
    from ws3d_b200 import iou3d_utils
    g = np.load(GOLD)
    a, b = torch.from_numpy(g["f2_boxes_a"]).to(dev), torch.from_numpy(g["f2_boxes_b"]).to(dev)
    i2, i3 = iou3d_utils.boxes_iou3d_aligned(a, b)
    # the fixture's BEV overlap came from host libm trig, the GPU uses libdevice: 1e-5
    np.testing.assert_allclose(i2.cpu().numpy(), g["f2_iou2d_diag"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(i3.cpu().numpy(), g["f2_iou3d_diag"], rtol=0, atol=1e-5)
    # bit-exact against the diagonal of the full fg x fg matrices (the idiom of train_functions.py:258-260)
    f2, f3 = iou3d_utils.boxes_iou3d_gpu(a, b)
    assert torch.equal(i2, torch.diagonal(f2)) and torch.equal(i3, torch.diagonal(f3))
    o2, o3 = oracle.boxes_iou3d_aligned(g["f2_boxes_a"], g["f2_boxes_b"])
    np.testing.assert_allclose(i3.cpu().numpy(), o3, rtol=0, atol=1e-5)
    np.testing.assert_allclose(i2.cpu().numpy(), o2, rtol=0, atol=1e-5)


def test_iou3d_aligned_edge_cases_and_large():
    from ws3d_b200 import iou3d_utils, synth
    e2, e3 = iou3d_utils.boxes_iou3d_aligned(torch.zeros(0, 7, device=dev), torch.zeros(0, 7, device=dev))
    assert e2.numel() == 0 and e3.numel() == 0
    pts = synth.make_scene(2)[:, :3]
    a = synth.make_boxes(pts, 16384, seed=5)
    rng = np.random.default_rng(3)
    b = a + rng.normal(0, 0.15, a.shape).astype(np.float32)
    b[:100] = a[:100]
    b[100:200, 3:6] = 0.0                      # degenerate boxes: clamp(min=1e-7) denominators
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    i2, i3 = iou3d_utils.boxes_iou3d_aligned(ta, tb)
    assert bool(torch.isfinite(i3).all()) and float(i3.min()) >= 0 and float(i3.max()) <= 1 + 1e-5
    assert float((i3[:100] - 1).abs().max()) < 1e-4      # self-IoU: the clipping's own rounding (same in the reference)
    # symmetric in its arguments up to the last ulp, and equal to the full matrix on a 1024-pair slice
    j2, j3 = iou3d_utils.boxes_iou3d_aligned(tb, ta)
    assert float((i3 - j3).abs().max()) < 1e-5
    f2, f3 = iou3d_utils.boxes_iou3d_gpu(ta[4000:5024], tb[4000:5024])
    assert torch.equal(i2[4000:5024], torch.diagonal(f2)) and torch.equal(i3[4000:5024], torch.diagonal(f3))
    with pytest.raises(RuntimeError):
        iou3d_utils.boxes_iou3d_aligned(ta[:4], tb[:5])


def _reference_radius_nms(centres, scores, radius=0.3):
    """tools/eval_auto.py:266-279, on the GPU in torch as the script runs it."""
    sort_points = torch.argsort(-scores)
    rois = centres[sort_points]
    keep_id = [0]
    d = _dist(rois, rois)
    for i in range(1, rois.shape[0]):
        if torch.min(d[keep_id, i], dim=-1)[0] > radius:
            keep_id.append(i)
    return sort_points[keep_id]


def test_radius_nms_matches_golden_oracle_and_reference_loop():
    from ws3d_b200 import native, proposal_utils
    g = np.load(GOLD)
    centres = torch.from_numpy(g["f3_centres"]).to(dev)
    scores = torch.from_numpy(g["f3_scores"]).to(dev)
    sorted_c = centres[torch.from_numpy(g["f3_sort"]).to(dev)].contiguous()
    keep, num = native.radius_nms_device(sorted_c, 0.3)
    got = keep[:int(num.item())].cpu().numpy()
    np.testing.assert_array_equal(got, g["f3_keep_id"])
    np.testing.assert_array_equal(got, oracle.radius_nms(sorted_c.cpu().numpy(), 0.3))
    # the whole wrapper (own argsort) against the script's loop run live; scores are distinct
    mine = proposal_utils.radius_nms(centres, scores, 0.3)
    want = _reference_radius_nms(centres, scores, 0.3)
    assert torch.equal(mine, want)


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 1000])
def test_radius_nms_small_and_ragged_sizes(n):
    from ws3d_b200 import native
    rng = np.random.default_rng(n)
    c = (rng.uniform(0, 3.0, (n, 2))).astype(np.float32)
    if n > 2:
        c[1] = c[0]                                       # duplicate of the best centre: suppressed
        c[2] = c[0] + np.array([0.3, 0.0], np.float32)    # around the threshold: decided by the float32 arithmetic
    keep, num = native.radius_nms_device(torch.from_numpy(c).to(dev), 0.3)
    np.testing.assert_array_equal(keep[:int(num.item())].cpu().numpy(), oracle.radius_nms(c, 0.3))


def test_radius_nms_empty_nan_and_large_invariants():
    from ws3d_b200 import native, synth
    keep, num = native.radius_nms_device(torch.zeros(0, 2, device=dev), 0.3)
    assert int(num.item()) == 0
    c = np.array([[0, 0], [np.nan, 1], [5, 5], [5.1, 5.0], [np.inf, 0]], np.float32)
    keep, num = native.radius_nms_device(torch.from_numpy(c).to(dev), 0.3)
    np.testing.assert_array_equal(keep[:int(num.item())].cpu().numpy(), oracle.radius_nms(c, 0.3))
    # 16384 candidates (BASELINE config-4 size): greedy invariants instead of an O(n^2) host loop
    pts = synth.make_scene(4)
    rng = np.random.default_rng(8)
    cen = (pts[rng.integers(0, 16384, 16384)][:, [0, 2]] + rng.normal(0, 0.2, (16384, 2))).astype(np.float32)
    t = torch.from_numpy(cen).to(dev)
    keep, num = native.radius_nms_device(t, 0.3)
    k = keep[:int(num.item())]
    assert 0 < k.numel() < 16384 and bool((k[1:] > k[:-1]).all()) and int(k[0]) == 0
    kept = t[k]
    d = _dist(kept, kept)
    d.fill_diagonal_(10.0)
    assert float(d.min()) > 0.3                                        # kept centres are pairwise farther than the radius
    dropped = torch.ones(16384, dtype=torch.bool, device=dev)
    dropped[k] = False
    di = torch.nonzero(dropped).flatten()
    dd = _dist(kept, t[di])                                            # (dropped, kept)
    earlier = k[None, :] < di[:, None]
    assert bool(((dd <= 0.3) & earlier).any(dim=1).all())              # each dropped one has an earlier kept centre within the radius
    np.testing.assert_array_equal(k.cpu().numpy(), oracle.radius_nms(cen, 0.3))


def test_cylinder_crop_matches_golden_oracle_and_dense_torch():
    from ws3d_b200 import proposal_utils
    g = np.load(GOLD)
    pts = torch.from_numpy(g["f3_points"]).to(dev)
    centres = torch.from_numpy(g["f3_centres"][g["f3_sort"]][g["f3_keep_id"]]).to(dev)
    cap = g["f3_idx"].shape[1]
    idx, cnt, any_ = proposal_utils.cylinder_crop(pts, centres, 4.0, cap=cap)
    np.testing.assert_array_equal(cnt.cpu().numpy(), g["f3_cnt"])
    np.testing.assert_array_equal(idx.cpu().numpy(), g["f3_idx"])
    np.testing.assert_array_equal(any_.cpu().numpy().astype(np.uint8), g["f3_any"])
    # live: the dense distance-matrix formulation of eval_auto.py:289-291,:336 on the GPU
    d = _dist(centres, pts[:, [0, 2]])
    assert torch.equal(any_, torch.min(d, dim=-1)[0] < 4.0)
    member = d < 4.0
    assert torch.equal(cnt.long(), member.sum(0))
    for c in (0, 7, centres.shape[0] - 1):
        want = torch.nonzero(member[:, c]).flatten()
        assert torch.equal(idx[c, :int(cnt[c])].long(), want)
    # a cap smaller than the membership keeps the first members and still counts all of them
    idx2, cnt2, _ = proposal_utils.cylinder_crop(pts, centres, 4.0, cap=16)
    assert torch.equal(cnt2, cnt) and torch.equal(idx2, idx[:, :16])


def test_cylinder_crop_edge_cases():
    from ws3d_b200 import proposal_utils
    pts = torch.tensor([[0, 0, 0], [4, 9, 0], [0, -3, 3.9999], [float("nan"), 0, 0], [100, 0, 100]], device=dev)
    cen = torch.tensor([[0.0, 0.0], [100.0, 100.0], [50.0, 50.0]], device=dev)
    idx, cnt, any_ = proposal_utils.cylinder_crop(pts, cen, 4.0)
    o_idx, o_cnt, o_any = oracle.cylinder_query(pts.cpu().numpy(), cen.cpu().numpy(), 4.0, 5)
    np.testing.assert_array_equal(idx.cpu().numpy(), o_idx)
    np.testing.assert_array_equal(cnt.cpu().numpy(), o_cnt)        # [2, 1, 0]: distance exactly 4 is outside, NaN is outside
    np.testing.assert_array_equal(any_.cpu().numpy(), o_any.astype(bool))
    assert cnt.tolist() == [2, 1, 0]
    idx, cnt, any_ = proposal_utils.cylinder_crop(pts, cen[:0], 4.0)
    assert idx.shape == (0, 5) and not bool(any_.any())
    idx, cnt, any_ = proposal_utils.cylinder_crop(pts[:0], cen, 4.0, cap=4)
    assert cnt.tolist() == [0, 0, 0]


def test_gaussian_labels_match_golden_and_oracle():
    """Row f4: ws3d_gaussian_rpn_labels against the reference's generate_gaussian_training_labels (fixture) and the
    oracle: reg_label bit-exact, cls_label to float32 rounding (float64 in the reference); batched with padded boxes."""
    from ws3d_b200 import label_utils, synth
    g = np.load(GOLD)
    pts, boxes = torch.from_numpy(g["f4_points"]).to(dev), torch.from_numpy(g["f4_boxes"]).to(dev)
    cls, reg = label_utils.generate_gaussian_training_labels(pts, boxes)
    np.testing.assert_array_equal(reg.cpu().numpy(), g["f4_reg"])
    np.testing.assert_allclose(cls.cpu().numpy(), g["f4_cls"], rtol=0, atol=1e-6)
    cls0, reg0 = label_utils.generate_gaussian_training_labels(pts[:100], boxes[:0])
    assert float(cls0.abs().max()) == 0.0 and float(reg0.abs().max()) == 0.0
    # a batch of three scenes with different numbers of (padded) boxes, n not a multiple of the block size
    rng = np.random.default_rng(4)
    scenes = np.stack([synth.make_scene(30 + k)[:5000, :3] for k in range(3)])
    counts = [9, 0, 23]
    padded = np.zeros((3, 23, 7), np.float32)
    for k, c in enumerate(counts):
        padded[k, :c] = synth.make_boxes(scenes[k], c, seed=k) if c else 0
        padded[k, c:] = rng.normal(0, 50, (23 - c, 7))                       # garbage in the padding must be ignored
    cls, reg = label_utils.generate_gaussian_training_labels(torch.from_numpy(scenes).to(dev), torch.from_numpy(padded).to(dev),
                                                             num_gt=torch.tensor(counts))
    for k, c in enumerate(counts):
        ocls, oreg = oracle.gaussian_rpn_labels(scenes[k], padded[k, :c])
        np.testing.assert_array_equal(reg[k].cpu().numpy(), oreg)
        np.testing.assert_allclose(cls[k].cpu().numpy(), ocls, rtol=0, atol=1e-6)
    assert float(cls[1].abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        label_utils.generate_gaussian_training_labels(torch.from_numpy(scenes).to(dev), torch.zeros(3, 600, 7, device=dev))
