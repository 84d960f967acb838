"""Golden vectors for the training-side rows (SURVEY.md section 8 f2 corner-loss box math, f4 subsampling, and the
Stage-1 loss of config 3), produced IN THE AUTHORING CONTAINER by executing the reference's own Python, read from
/root/reference at generation time (nothing is copied into this repo):

  corners   lib/utils/kitti_utils.py::boxes3d_to_corners3d_torch, imported unmodified (torch.cuda.FloatTensor is aliased
            to the CPU type: this container has no GPU; same dtype, same arithmetic up to the matmul's summation order).
  corner    the corner-distance lines of lib/net/train_functions.py:266-271 driven by that function, and the gradient
            torch's autograd gives for sum(dist * g) with respect to the predicted boxes.
  rpn_loss  lib/utils/loss_utils.py::SigmoidFocalClassificationLoss and ::get_rpn_reg_loss, imported unmodified (its
            `import lib.utils.iou3d.iou3d_utils` is satisfied with a stub of the native module), driven by the
            SigmoidFocalLoss / Gaussian_Center branch of lib/net/train_functions.py:160-228.
  subsample the block lib/datasets/kitti_rcnn_dataset.py:424-444 cut out of the (un-importable) module's source text
            and executed as it stands after np.random.seed(...).

    python tools/make_golden_train.py      ->  tests/golden/train_rows.npz
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs  # noqa: E402
from ws3d_b200 import synth  # noqa: E402


def load_reference():
    for name in ("iou3d_cuda", "roipool3d_cuda", "pointnet2_cuda"):
        sys.modules.setdefault(name, types.ModuleType(name))     # native modules: not called by anything used here
    sys.path.insert(0, REF)
    torch.cuda.FloatTensor = torch.FloatTensor
    import lib.utils.kitti_utils as ku
    import lib.utils.loss_utils as lu
    return ku, lu


def reference_subsample(pts_rect, pts_intensity, pts_rect_depth, npoints, seed):
    src = open(os.path.join(REF, "lib/datasets/kitti_rcnn_dataset.py")).read().splitlines()
    start = next(i for i, l in enumerate(src) if "if self.npoints < len(pts_rect):" in l)
    end = next(i for i, l in enumerate(src) if "ret_pts_intensity = pts_intensity[choice] - 0.5" in l)
    block = textwrap.dedent("\n".join(src[start:end + 1])).replace("self.npoints", "npoints")
    ns = {"np": np, "pts_rect": pts_rect, "pts_intensity": pts_intensity, "pts_depth": pts_rect_depth, "npoints": npoints}
    np.random.seed(seed)
    exec(block, ns)
    return ns["ret_pts_rect"], ns["ret_pts_intensity"], ns["choice"]


def main():
    ku, lu = load_reference()
    out = {}
    rng = np.random.default_rng(99)
    # ---- corners / corner distance
    scene = synth.make_scene(0)
    gt = synth.make_boxes(scene[:, :3], 257, seed=11)
    pred = gt + rng.normal(0, 0.15, gt.shape).astype(np.float32)
    pred[5] = gt[5]                      # an exact match (distance 0: the norm's subgradient)
    pred[6, 6] = gt[6, 6] + np.float32(np.pi)   # the flipped ground truth is the nearer one
    out["boxes_gt"], out["boxes_pred"] = gt, pred
    for flip in (False, True):
        out[f"corners_flip{int(flip)}"] = ku.boxes3d_to_corners3d_torch(torch.from_numpy(gt), flip=flip).numpy()
    p = torch.from_numpy(pred).clone().requires_grad_(True)
    g = torch.from_numpy(gt)
    gt_fcorner = g.clone()
    pred_corner = ku.boxes3d_to_corners3d_torch(p)                       # train_functions.py:266
    gt_corner = ku.boxes3d_to_corners3d_torch(gt_fcorner)                # :267
    gt_fcorner[:, 6] += np.pi                                            # :268
    gt_flip_corner = ku.boxes3d_to_corners3d_torch(gt_fcorner)           # :269
    corner_dist = torch.min(torch.norm(pred_corner - gt_corner, dim=-1),
                            torch.norm(pred_corner - gt_flip_corner, dim=-1))   # :270-271
    corner_loss = F.smooth_l1_loss(corner_dist, torch.zeros_like(corner_dist))  # :272-273
    gw = torch.from_numpy(rng.normal(0, 1, corner_dist.shape).astype(np.float32))
    (corner_dist * gw).sum().backward()
    out["corner_dist"], out["corner_loss"] = corner_dist.detach().numpy(), np.float32(corner_loss.item())
    out["corner_grad_w"], out["corner_grad_pred"] = gw.numpy(), p.grad.numpy()
    # ---- Stage-1 loss
    B, N = 2, 4096
    rpn_cls_np, rpn_reg_np, label, reg_label = golden_inputs.rpn_loss_inputs(B, N)
    rpn_cls, rpn_reg = torch.from_numpy(rpn_cls_np), torch.from_numpy(rpn_reg_np)
    rpn_cls_label, rpn_reg_label = torch.from_numpy(label), torch.from_numpy(reg_label)
    loss_func = lu.SigmoidFocalClassificationLoss(alpha=0.25, gamma=2.0)          # rpn.py:51-52
    flat, cls_flat = rpn_cls_label.view(-1), rpn_cls.view(-1)                     # train_functions.py:166-167
    fg_mask = flat > 0                                                            # :169
    rpn_cls_target, pos, neg = flat.float(), flat.float(), (1 - flat).float()     # :177-179
    cls_weights = (pos + neg) / torch.clamp(pos.sum(), min=1.0)                   # :185-187
    per = loss_func(cls_flat, rpn_cls_target, cls_weights)                        # :188
    point_num = B * N
    loss_loc, d = lu.get_rpn_reg_loss(rpn_reg.view(point_num, -1)[fg_mask], rpn_reg_label.view(point_num, 3)[fg_mask],
                                      loc_scope=4.0, loc_bin_size=0.8)            # :211-215
    out.update({"rpn_loss_cls": np.float32(per.sum().item()), "rpn_loss_cls_pos": np.float32((per * pos).sum().item()),
                "rpn_loss_cls_neg": np.float32((per * neg).sum().item()), "rpn_loss_reg": np.float32(loss_loc.item()),
                "rpn_loss": np.float32((per.sum() + loss_loc).item()), "rpn_fg_sum": np.int64(fg_mask.long().sum().item())})
    for k, v in d.items():
        out["rpn_" + k] = np.float32(v)
    # ---- subsampling: more points than npoints (near / far split), fewer (tiling), and equal
    for tag, n, npoints, seed in golden_inputs.SUBSAMPLE_CASES:
        pts_rect, depth, inten = golden_inputs.subsample_inputs(tag, n, seed)
        ret_rect, ret_int, choice = reference_subsample(pts_rect, inten, depth, npoints, seed)
        assert np.array_equal(ret_rect, pts_rect[choice]) and np.array_equal(ret_int, inten[choice] - 0.5)
        out[f"sub_{tag}_choice"] = choice.astype(np.int32)
        out[f"sub_{tag}_intensity_head"] = ret_int[:64].astype(np.float32)
    path = os.path.join(ROOT, "tests", "golden", "train_rows.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
