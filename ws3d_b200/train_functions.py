"""Stage-1 RPN training loss, restated from the reference without host synchronisation.

`get_rpn_loss` follows lib/net/train_functions.py:160-228 for the configuration the WS3D Stage-1 recipe trains with
(tools/cfgs/weaklyRPN.yaml:31,59-64: Gaussian_Center = True, LOSS_CLS = SigmoidFocalLoss, FOCAL_ALPHA[0] = 0.25,
FOCAL_GAMMA = 2.0, LOSS_WEIGHT = [1, 1]) and `get_rpn_reg_loss` follows lib/utils/loss_utils.py:88-148.  The reference
selects the foreground rows with a boolean mask (`pred[fg_mask]`, a host round trip for the output size) and calls
`.item()` on every term; here the same sums run over all points with the mask as a weight and the foreground count
clamped to >= 1, so the step has static shapes and no synchronisation (it replays as one CUDA graph).  Values agree with
the reference's to summation order (tests/test_train_functions.py checks against the reference's own functions).
"""
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import native

FOCAL_ALPHA, FOCAL_GAMMA = 0.25, 2.0          # weaklyRPN.yaml:61-62
LOSS_WEIGHT = (1.0, 1.0)                      # weaklyRPN.yaml:64
LOC_SCOPE, LOC_BIN_SIZE = 4.0, 0.8            # weaklyRPN.yaml:37-38


def sigmoid_focal_loss(logits: torch.Tensor, target: torch.Tensor, weights: torch.Tensor,
                       alpha: float = FOCAL_ALPHA, gamma: float = FOCAL_GAMMA) -> torch.Tensor:
    """SigmoidFocalClassificationLoss.forward (loss_utils.py:42-74) with soft targets, elementwise."""
    ce = torch.clamp(logits, min=0) - logits * target + torch.log1p(torch.exp(-torch.abs(logits)))   # :77-85
    p = torch.sigmoid(logits)
    p_t = target * p + (1 - target) * (1 - p)
    mod = torch.pow(1.0 - p_t, gamma) if gamma else 1.0
    alpha_w = target * alpha + (1 - target) * (1 - alpha)
    return mod * alpha_w * ce * weights


def get_rpn_reg_loss(pred_reg: torch.Tensor, reg_label: torch.Tensor, fg: torch.Tensor, loc_scope: float = LOC_SCOPE,
                     loc_bin_size: float = LOC_BIN_SIZE) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """loss_utils.get_rpn_reg_loss (:88-148) over ALL rows with `fg` (P,) in {0, 1} as the row weight: the reference's
    means over the selected rows become weighted sums divided by the foreground count.
    pred_reg (P, 4 * per_loc_bin_num), reg_label (P, 3) [dx, 0, dz]."""
    nbin = int((loc_scope + 1e-3) / loc_bin_size) * 2
    assert pred_reg.shape[1] == 4 * nbin, '%d vs %d' % (pred_reg.shape[1], 4 * nbin)
    count = fg.sum().clamp(min=1.0)
    x_shift = torch.clamp(reg_label[:, 0] + loc_scope, 0, loc_scope * 2 - 1e-3)
    z_shift = torch.clamp(reg_label[:, 2] + loc_scope, 0, loc_scope * 2 - 1e-3)
    x_bin = (x_shift / loc_bin_size).floor().long()
    z_bin = (z_shift / loc_bin_size).floor().long()
    loss_x_bin = (F.cross_entropy(pred_reg[:, 0:nbin], x_bin, reduction='none') * fg).sum() / count
    loss_z_bin = (F.cross_entropy(pred_reg[:, nbin:2 * nbin], z_bin, reduction='none') * fg).sum() / count
    x_res = (x_shift - (x_bin.float() * loc_bin_size + loc_bin_size / 2)) / (loc_bin_size / 2)
    z_res = (z_shift - (z_bin.float() * loc_bin_size + loc_bin_size / 2)) / (loc_bin_size / 2)
    x_pred = pred_reg[:, 2 * nbin:3 * nbin].gather(1, x_bin[:, None]).squeeze(1)     # (pred * onehot).sum(dim=1)
    z_pred = pred_reg[:, 3 * nbin:4 * nbin].gather(1, z_bin[:, None]).squeeze(1)
    loss_x_res = (F.smooth_l1_loss(x_pred, x_res, reduction='none') * fg).sum() / count
    loss_z_res = (F.smooth_l1_loss(z_pred, z_res, reduction='none') * fg).sum() / count
    terms = {'loss_x_bin': loss_x_bin, 'loss_z_bin': loss_z_bin, 'loss_x_res': loss_x_res, 'loss_z_res': loss_z_res}
    return loss_x_bin + loss_z_bin + loss_x_res + loss_z_res, terms


def get_rpn_loss(rpn_cls: torch.Tensor, rpn_reg: torch.Tensor, rpn_cls_label: torch.Tensor, rpn_reg_label: torch.Tensor
                 ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """train_functions.get_rpn_loss (:160-228), SigmoidFocalLoss / Gaussian_Center branch.
    rpn_cls (B,N,1), rpn_reg (B,N,C), rpn_cls_label (B,N) soft labels in [0,1], rpn_reg_label (B,N,3).
    Returns (loss, tensors for logging) -- nothing is copied to the host."""
    label = rpn_cls_label.reshape(-1).float()
    logit = rpn_cls.reshape(-1)
    fg = (label > 0).float()
    pos, neg = label, 1 - label                                   # Gaussian_Center: soft positives / negatives
    weights = (pos + neg) / torch.clamp(pos.sum(), min=1.0)
    per_point = sigmoid_focal_loss(logit, label, weights)
    loss_cls = per_point.sum()
    points = rpn_reg.shape[0] * rpn_reg.shape[1]
    loss_reg, terms = get_rpn_reg_loss(rpn_reg.reshape(points, -1), rpn_reg_label.reshape(points, 3), fg)
    loss = loss_cls * LOSS_WEIGHT[0] + loss_reg * LOSS_WEIGHT[1]
    terms.update({'rpn_loss_cls': loss_cls, 'rpn_loss_reg': loss_reg, 'rpn_loss': loss, 'rpn_fg_sum': fg.sum(),
                  'rpn_loss_cls_pos': (per_point * pos).sum(), 'rpn_loss_cls_neg': (per_point * neg).sum()})
    return loss, terms


class _CornerDistance(torch.autograd.Function):
    """dist (N,8) = min(|P - G|, |P - G_flipped|) per corner for aligned (N,7) boxes (train_functions.py:266-271); the
    gradient flows to the predicted boxes only, as in the reference (its ground truth is a detached clone)."""

    @staticmethod
    def forward(ctx, pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
        pred, gt = pred.contiguous().float(), gt.detach().contiguous().float()
        dist = torch.empty((pred.shape[0], 8), dtype=torch.float32, device=pred.device)
        native.corner_distance(pred, gt, dist)
        ctx.save_for_backward(pred, gt)
        return dist

    @staticmethod
    def backward(ctx, grad_dist):
        pred, gt = ctx.saved_tensors
        grad_pred = torch.empty_like(pred)
        native.corner_distance_grad(pred, gt, grad_dist.contiguous().float(), grad_pred)
        return grad_pred, None


corner_distance = _CornerDistance.apply


def corner_loss(pred_boxes3d: torch.Tensor, gt_boxes3d: torch.Tensor) -> torch.Tensor:
    """The corner loss of the Stage-2 training step (train_functions.py:264-273): smooth-L1 of the corner distance
    against zero, mean over (boxes, corners).  Three corner computations, two norms, a min: one launch each way."""
    dist = corner_distance(pred_boxes3d, gt_boxes3d)
    return F.smooth_l1_loss(dist, torch.zeros_like(dist))
