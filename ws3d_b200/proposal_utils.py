"""Proposal clustering between the two WS3D stages.

The reference's live inference script (tools/eval_auto.py) does not use box NMS / roipool3d between its
stages: it thins the predicted object centres with a greedy *radius* NMS in the BEV plane (:263-279, a Python loop
over a P x P distance matrix) and crops a 4 m cylinder of points round every surviving centre (:289-291, :327-343,
a P x N distance matrix plus one boolean-mask gather per centre).  These are the same two steps on libws3d_ops.so;
results are bit-identical (same float32 distance arithmetic as lib/utils/distance.py:3, same comparisons).
"""
from typing import Optional, Tuple

import torch

from . import native


def radius_nms(centers_xz: torch.Tensor, scores: torch.Tensor, radius: float = 0.3) -> torch.Tensor:
    """centers_xz (P,2) CUDA float32, scores (P).  Returns the indices (into the input) of the kept centres, best
    score first -- `sort_points[keep_id]` of eval_auto.py:266-279: centre i (in descending-score order) is kept iff
    it is farther than `radius` from every centre kept before it."""
    order = torch.argsort(-scores)
    sorted_centers = centers_xz[order].contiguous().float()
    keep, num = native.radius_nms_device(sorted_centers, radius)
    return order[keep[:int(num.item())]].contiguous()


def cylinder_crop(points: torch.Tensor, centers_xz: torch.Tensor, radius: float = 4.0,
                  cap: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """points (N, >=3) CUDA float32 [x,y,z,...], centers_xz (M,2).  Returns
         idx (M,cap) int32 : the members of every cylinder in point-index order, -1 padded
                             (row c == nonzero(distance_2(centers, points[:, [0, 2]])[:, c] < radius), eval_auto.py:336),
         cnt (M)     int32 : the member counts (may exceed cap; cap defaults to N so nothing is cut),
         any (N)     bool  : member of at least one cylinder (`cur_proposal_points_index`, eval_auto.py:291).
    The per-proposal network input of eval_auto.py:338-343 is then points[idx[c, :cnt[c]]] shifted by the centre."""
    n, m = points.shape[0], centers_xz.shape[0]
    cap = n if cap is None else int(cap)
    xyz = points[:, :3].contiguous().float()
    idx = torch.full((m, cap), -1, dtype=torch.int32, device=points.device)
    cnt = torch.zeros(m, dtype=torch.int32, device=points.device)
    any_flag = torch.zeros(n, dtype=torch.uint8, device=points.device)
    if m:
        native.cylinder_query(xyz, centers_xz.contiguous().float(), radius, idx, cnt, any_flag)
    return idx, cnt, any_flag.bool()
