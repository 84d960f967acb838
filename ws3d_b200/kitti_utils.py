"""The box helpers of lib/utils/kitti_utils.py that the hot-path wrappers and the loss-side box math depend on."""
import numpy as np
import torch


def boxes3d_to_bev_torch(boxes3d: torch.Tensor) -> torch.Tensor:
    """(N,7) [x,y,z,h,w,l,ry] -> (N,5) [x1,y1,x2,y2,ry] in the x-z plane (kitti_utils.py:134-147)."""
    cu, cv = boxes3d[:, 0], boxes3d[:, 2]
    half_l, half_w = boxes3d[:, 5] / 2, boxes3d[:, 4] / 2
    return torch.stack((cu - half_l, cv - half_w, cu + half_l, cv + half_w, boxes3d[:, 6]), dim=1)


def enlarge_box3d(boxes3d, extra_width):
    """Grow h, w, l by 2*extra_width and move the bottom centre down by extra_width (kitti_utils.py:150-160)."""
    large = boxes3d.copy() if isinstance(boxes3d, np.ndarray) else boxes3d.clone()
    large[:, 3:6] += extra_width * 2
    large[:, 1] += extra_width
    return large


def boxes3d_to_corners3d_torch(boxes3d: torch.Tensor, flip: bool = False) -> torch.Tensor:
    """(N,7) [x,y,z,h,w,l,ry] -> (N,8,3) rotated corners (kitti_utils.py:104-131); one launch instead of ~15 eager
    kernels.  Forward only, like every use of it outside the corner loss (train_functions.corner_distance has the
    differentiable form)."""
    from . import native
    boxes = boxes3d.detach().contiguous().float()
    corners = torch.empty((boxes.shape[0], 8, 3), dtype=torch.float32, device=boxes.device)
    native.boxes3d_to_corners3d(boxes, flip, corners)
    return corners
