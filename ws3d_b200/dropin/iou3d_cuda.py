"""Drop-in for the reference extension module `iou3d_cuda` (lib/utils/iou3d/src/iou3d.cpp:175-178)."""
from ws3d_b200.native import boxes_iou_bev_gpu, boxes_overlap_bev_gpu, nms_gpu, nms_normal_gpu  # noqa: F401
