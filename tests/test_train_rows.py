"""Training-side rows: the Stage-1 loss restatement, the corner-loss box math and the loader
subsampling (f4), against tests/golden/train_rows.npz -- outputs of the reference's own Python
(tools/make_golden_train.py).  CPU tests pin the oracle restatements and the host logic; `-m gpu` tests pin the
kernels through the C ABI."""
import os

import numpy as np
import pytest
import torch

import golden_inputs
import oracle
from ws3d_b200 import data_utils, train_functions

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_rows.npz"))


# ---- CPU -----------------------------------------------------------------------------------------
def test_rpn_loss_restatement_matches_reference_functions():
    rpn_cls, rpn_reg, label, reg_label = (torch.from_numpy(a) for a in golden_inputs.rpn_loss_inputs())
    loss, terms = train_functions.get_rpn_loss(rpn_cls, rpn_reg, label, reg_label)
    for key in ("rpn_loss_cls", "rpn_loss_cls_pos", "rpn_loss_cls_neg", "rpn_loss_reg", "rpn_loss", "loss_x_bin", "loss_z_bin",
                "loss_x_res", "loss_z_res"):
        want = float(GOLD[key if key.startswith("rpn_") else "rpn_" + key])
        assert abs(float(terms[key]) - want) <= 2e-6 * max(1.0, abs(want)), (key, float(terms[key]), want)
    assert int(terms["rpn_fg_sum"]) == int(GOLD["rpn_fg_sum"])
    assert abs(float(loss) - float(GOLD["rpn_loss"])) <= 2e-6 * abs(float(GOLD["rpn_loss"]))


def test_rpn_loss_has_no_foreground_branch():
    """fg_sum == 0: the reference returns rpn_loss_cls * 0 for the regression term (train_functions.py:216-217)."""
    cls, reg = torch.randn(1, 64, 1), torch.randn(1, 64, 40)
    loss, terms = train_functions.get_rpn_loss(cls, reg, torch.zeros(1, 64), torch.zeros(1, 64, 3))
    assert float(terms["rpn_loss_reg"]) == 0.0 and torch.isfinite(loss)


def test_rpn_loss_gradient_matches_masked_reference_formulation():
    """The weighted-sum form has the same gradient as the reference's boolean-mask form."""
    import torch.nn.functional as F
    rpn_cls, rpn_reg, label, reg_label = (torch.from_numpy(a) for a in golden_inputs.rpn_loss_inputs(1, 512))
    reg_a = rpn_reg.clone().requires_grad_(True)
    _, terms = train_functions.get_rpn_loss(rpn_cls, reg_a, label, reg_label)
    terms["rpn_loss_reg"].backward()
    reg_b = rpn_reg.clone().requires_grad_(True)
    fg = label.view(-1) > 0
    pred, lab = reg_b.view(512, -1)[fg], reg_label.view(512, 3)[fg]
    xs, zs = torch.clamp(lab[:, 0] + 4.0, 0, 8 - 1e-3), torch.clamp(lab[:, 2] + 4.0, 0, 8 - 1e-3)
    xb, zb = (xs / 0.8).floor().long(), (zs / 0.8).floor().long()
    loss = F.cross_entropy(pred[:, 0:10], xb) + F.cross_entropy(pred[:, 10:20], zb)
    loss = loss + F.smooth_l1_loss(pred[:, 20:30].gather(1, xb[:, None]).squeeze(1), (xs - (xb.float() * 0.8 + 0.4)) / 0.4)
    loss = loss + F.smooth_l1_loss(pred[:, 30:40].gather(1, zb[:, None]).squeeze(1), (zs - (zb.float() * 0.8 + 0.4)) / 0.4)
    loss.backward()
    torch.testing.assert_close(reg_a.grad, reg_b.grad, rtol=1e-5, atol=1e-7)


def test_oracle_corners_match_reference_python():
    for flip in (0, 1):
        got = oracle.boxes3d_to_corners3d(GOLD["boxes_gt"], bool(flip))
        np.testing.assert_allclose(got, GOLD[f"corners_flip{flip}"], rtol=0, atol=2e-5)   # matmul summation order, libm trig


def test_oracle_corner_distance_matches_reference_python():
    got = oracle.corner_distance(GOLD["boxes_pred"], GOLD["boxes_gt"])
    np.testing.assert_allclose(got, GOLD["corner_dist"], rtol=1e-4, atol=3e-5)
    assert np.all(got[5] == 0)                                  # the exact match
    loss = torch.nn.functional.smooth_l1_loss(torch.from_numpy(got), torch.zeros(got.shape))
    assert abs(float(loss) - float(GOLD["corner_loss"])) < 1e-5


@pytest.mark.parametrize("tag,n,npoints,seed", golden_inputs.SUBSAMPLE_CASES)
def test_subsample_restatement_and_host_draws_match_reference_block(tag, n, npoints, seed):
    pts_rect, depth, inten = golden_inputs.subsample_inputs(tag, n, seed)
    pts = np.concatenate([pts_rect, inten[:, None]], axis=1)
    out, choice = oracle.subsample_points(pts, depth, npoints, np.random.RandomState(seed))
    np.testing.assert_array_equal(choice, GOLD[f"sub_{tag}_choice"])
    np.testing.assert_array_equal(out[:64, 3], GOLD[f"sub_{tag}_intensity_head"])
    # the draws handed to the kernel select the same rows (numpy emulation of the kernel's indexing)
    n_near = int((depth < 40.0).sum())
    perm, order = data_utils.draw_subsample(np.random.RandomState(seed), n, n_near, npoints)
    if n > npoints:
        near, far = np.where(depth < 40.0)[0], np.where(~(depth < 40.0))[0]
        sel = np.concatenate([near[perm], far])
    else:
        sel = perm % n
    np.testing.assert_array_equal(sel[order], choice)


def test_draw_subsample_rejects_what_numpy_rejects():
    with pytest.raises(ValueError):
        data_utils.draw_subsample(np.random.RandomState(0), 100, 10, 50)   # 90 far points > npoints


# ---- GPU -----------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_corners_kernel_matches_reference_and_oracle():
    from ws3d_b200 import kitti_utils
    boxes = torch.from_numpy(GOLD["boxes_gt"]).cuda()
    for flip in (False, True):
        got = kitti_utils.boxes3d_to_corners3d_torch(boxes, flip).cpu().numpy()
        np.testing.assert_allclose(got, GOLD[f"corners_flip{int(flip)}"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(got, oracle.boxes3d_to_corners3d(GOLD["boxes_gt"], flip), rtol=0, atol=2e-5)
    # against the reference's formula evaluated by torch on the same GPU
    b = boxes
    h, w, l, ry = b[:, 3:4], b[:, 4:5], b[:, 5:6], b[:, 6:7]
    zeros, ones = torch.zeros_like(h), torch.ones_like(h)
    xc = torch.cat([l / 2., l / 2., -l / 2., -l / 2., l / 2., l / 2., -l / 2., -l / 2.], dim=1)
    yc = torch.cat([zeros, zeros, zeros, zeros, -h, -h, -h, -h], dim=1)
    zc = torch.cat([w / 2., -w / 2., -w / 2., w / 2., w / 2., -w / 2., -w / 2., w / 2.], dim=1)
    corners = torch.stack([xc, yc, zc], dim=1)
    cosa, sina = torch.cos(ry), torch.sin(ry)
    R = torch.stack([torch.cat([cosa, zeros, sina], 1), torch.cat([zeros, ones, zeros], 1), torch.cat([-sina, zeros, cosa], 1)], 1)
    ref = (torch.matmul(R, corners) + b[:, 0:3].unsqueeze(2)).permute(0, 2, 1)
    torch.testing.assert_close(kitti_utils.boxes3d_to_corners3d_torch(boxes), ref, rtol=0, atol=2e-5)


@pytest.mark.gpu
def test_corner_distance_forward_and_gradient():
    pred = torch.from_numpy(GOLD["boxes_pred"]).cuda().requires_grad_(True)
    gt = torch.from_numpy(GOLD["boxes_gt"]).cuda()
    dist = train_functions.corner_distance(pred, gt)
    np.testing.assert_allclose(dist.detach().cpu().numpy(), GOLD["corner_dist"], rtol=1e-4, atol=3e-5)
    np.testing.assert_allclose(dist.detach().cpu().numpy(), oracle.corner_distance(GOLD["boxes_pred"], GOLD["boxes_gt"]), rtol=1e-4, atol=3e-5)
    (dist * torch.from_numpy(GOLD["corner_grad_w"]).cuda()).sum().backward()
    got, want = pred.grad.cpu().numpy(), GOLD["corner_grad_pred"]
    # rows 5 (distance 0) and 6 (flip tie region) included; the gradient is O(10) for l, w, ry: relative tolerance
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-4)
    loss = train_functions.corner_loss(pred.detach(), gt)
    assert abs(float(loss) - float(GOLD["corner_loss"])) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n,npoints,seed", golden_inputs.SUBSAMPLE_CASES)
def test_subsample_kernel_equals_reference_block(tag, n, npoints, seed):
    pts_rect, depth, inten = golden_inputs.subsample_inputs(tag, n, seed)
    pts = torch.from_numpy(np.concatenate([pts_rect, inten[:, None]], axis=1)).cuda()
    n_near = int((depth < 40.0).sum())
    perm, order = data_utils.draw_subsample(np.random.RandomState(seed), n, n_near, npoints)
    out, choice, status = data_utils.subsample_points(pts, torch.from_numpy(depth).cuda(), npoints, perm, order, n_near)
    assert int(status.item()) == 0
    np.testing.assert_array_equal(choice.cpu().numpy(), GOLD[f"sub_{tag}_choice"])
    want, _ = oracle.subsample_points(pts.cpu().numpy(), depth, npoints, np.random.RandomState(seed))
    np.testing.assert_array_equal(out.cpu().numpy(), want)       # rows are copies; intensity - 0.5 is one exact-rounded op


@pytest.mark.gpu
def test_subsample_kernel_flags_a_wrong_near_count():
    pts_rect, depth, inten = golden_inputs.subsample_inputs("down", 21000, 3)
    pts = torch.from_numpy(np.concatenate([pts_rect, inten[:, None]], axis=1)).cuda()
    n_near = int((depth < 40.0).sum()) - 1
    perm, order = data_utils.draw_subsample(np.random.RandomState(0), 21000, n_near, 16384)
    _, _, status = data_utils.subsample_points(pts, torch.from_numpy(depth).cuda(), 16384, perm, order, n_near)
    assert int(status.item()) == 1


@pytest.mark.gpu
def test_sample_objects_equals_per_object_calls_without_syncs():
    from ws3d_b200 import pointnet2_utils
    rng = np.random.default_rng(0)
    objs = [torch.from_numpy(rng.normal(0, 1, (n, 3)).astype(np.float32)).cuda() for n in (137, 100, 512, 260)]
    got = data_utils.sample_objects(objs, 100)
    for o, g in zip(objs, got):
        want = oracle.furthest_point_sample(o.cpu().numpy()[None], 100)[0]
        np.testing.assert_array_equal(g.cpu().numpy(), want)
