"""Stage-1 training labels on the GPU.

The reference builds its RPN targets per sample in numpy inside the data loader, which runs with num_workers = 0:
`KittiRCNNDataset.generate_gaussian_training_labels` (lib/datasets/kitti_rcnn_dataset.py:529-573).  This is the same
computation for a whole batch in one launch; reg_label is bit-identical, cls_label (a float64 scipy Gaussian there)
agrees to float32 rounding.  The other two data-path steps named by that row need nothing new: the GT-paste sampler
(:309-311) is `pointnet2_utils.furthest_point_sample` at B = 1, and the 16384-point subsampling (:424-452) is a random
choice whose RNG stream (numpy's) a GPU version would not reproduce.
"""
from typing import Optional, Tuple

import torch

from . import native

GAUSS_HEIGHT, GAUSS_STATUS, GAUSS_COV = 0.707, 0.7, 1.5   # lib/config.py:45-47 = tools/cfgs/weaklyRPN.yaml:32-34
FG_RADIUS = 4.0                                          # kitti_rcnn_dataset.py:568


def generate_gaussian_training_labels(pts_rect: torch.Tensor, gt_boxes3d: torch.Tensor, num_gt: Optional[torch.Tensor] = None,
                                      gauss_height: float = GAUSS_HEIGHT, gauss_status: float = GAUSS_STATUS,
                                      gauss_cov: float = GAUSS_COV) -> Tuple[torch.Tensor, torch.Tensor]:
    """pts_rect (B,N,3) or (N,3) CUDA float32; gt_boxes3d (B,G,7) or (G,7) [x,y,z,h,w,l,ry] (padded rows ignored when
    `num_gt` (B) gives the valid count per scene).  Returns cls_label (B,N) and reg_label (B,N,3) (without the batch
    dimension for unbatched input), as the reference's static method does per sample."""
    single = pts_rect.dim() == 2
    pts = (pts_rect.unsqueeze(0) if single else pts_rect).contiguous().float()
    boxes = (gt_boxes3d.unsqueeze(0) if gt_boxes3d.dim() == 2 else gt_boxes3d).contiguous().float()
    cnt = None if num_gt is None else num_gt.to(device=pts.device, dtype=torch.int32).contiguous()
    cls_label = torch.empty(pts.shape[:2], dtype=torch.float32, device=pts.device)
    reg_label = torch.empty(pts.shape, dtype=torch.float32, device=pts.device)
    native.gaussian_rpn_labels(pts, boxes, cnt, gauss_height, gauss_status, gauss_cov, FG_RADIUS, cls_label, reg_label)
    return (cls_label[0], reg_label[0]) if single else (cls_label, reg_label)
