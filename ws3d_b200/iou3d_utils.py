"""Host-side mirror of lib/utils/iou3d/iou3d_utils.py: boxes_iou_bev, boxes_iou3d_gpu, nms_gpu,
nms_normal_gpu with the reference's arguments and return values, on libws3d_ops.so."""
import torch

from . import kitti_utils, native


def boxes_iou_bev(boxes_a: torch.Tensor, boxes_b: torch.Tensor) -> torch.Tensor:
    """(M,5), (N,5) [x1,y1,x2,y2,ry] -> rotated BEV IoU (M,N)   (iou3d_utils.py:6-18)."""
    ans = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    native.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans)
    return ans


def boxes_iou3d_gpu(boxes_a: torch.Tensor, boxes_b: torch.Tensor):
    """(N,7), (M,7) [x,y,z,h,w,l,ry] -> (iou2d, iou3d), both (N,M)   (iou3d_utils.py:21-56)."""
    a_bev = kitti_utils.boxes3d_to_bev_torch(boxes_a).contiguous()
    b_bev = kitti_utils.boxes3d_to_bev_torch(boxes_b).contiguous()
    overlaps_bev = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    native.boxes_overlap_bev_gpu(a_bev, b_bev, overlaps_bev)

    a_hmin, a_hmax = (boxes_a[:, 1] - boxes_a[:, 3]).view(-1, 1), boxes_a[:, 1].view(-1, 1)
    b_hmin, b_hmax = (boxes_b[:, 1] - boxes_b[:, 3]).view(1, -1), boxes_b[:, 1].view(1, -1)
    overlaps_h = torch.clamp(torch.min(a_hmax, b_hmax) - torch.max(a_hmin, b_hmin), min=0)

    s_a = (boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    s_b = (boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    iou2d = overlaps_bev / torch.clamp(s_a + s_b - overlaps_bev, min=1e-7)

    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    iou3d = overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-7)
    return iou2d, iou3d


def boxes_iou3d_aligned(boxes_a: torch.Tensor, boxes_b: torch.Tensor):
    """(n,7), (n,7) -> (iou2d (n), iou3d (n)) of box i of `a` with box i of `b`: the diagonal of boxes_iou3d_gpu in one
    launch and O(n) work.  Drop-in for the `torch.gather(iou3d, 1, eye)` idiom of lib/net/train_functions.py:258-260
    and :287-289, which builds the whole fg x fg matrix (twice per step) to keep n numbers; bit-identical to it."""
    a, b = boxes_a.contiguous().float(), boxes_b.contiguous().float()
    iou2d = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
    iou3d = torch.empty_like(iou2d)
    native.boxes_iou3d_aligned(a, b, iou2d, iou3d)
    return iou2d, iou3d


def _nms(boxes, scores, thresh, rotated):
    order = scores.sort(0, descending=True)[1]
    sorted_boxes = boxes[order].contiguous()
    keep, num = native.nms_device(sorted_boxes, thresh, rotated=rotated)
    return order[keep[:int(num.item())]].contiguous()


def nms_gpu(boxes: torch.Tensor, scores: torch.Tensor, thresh: float) -> torch.Tensor:
    """Rotated NMS; returns indices into `boxes` of the kept boxes, best first (iou3d_utils.py:59-73).
    The greedy scan runs on the device; only the kept count crosses to the host."""
    return _nms(boxes, scores, thresh, True)


def nms_normal_gpu(boxes: torch.Tensor, scores: torch.Tensor, thresh: float) -> torch.Tensor:
    """Axis-aligned NMS (iou3d_utils.py:76-90)."""
    return _nms(boxes, scores, thresh, False)
