"""ws3d_b200: the WS3D PointNet++ set-abstraction / roipool3d / iou3d hot path, hand-written for
B200 (sm_100a) behind a C ABI (include/ws3d_ops.h, libws3d_ops.so).

    from ws3d_b200 import pointnet2_utils, pointnet2_modules, iou3d_utils, roipool3d_utils

mirror the reference's Python op wrappers; `install_dropins()` registers `pointnet2_cuda`,
`iou3d_cuda` and `roipool3d_cuda` so the reference's own wrappers run on top unmodified.
There is no CPU or PyTorch fallback: a missing library raises at first use.
"""
import importlib
import os
import sys

__version__ = "0.1.0"

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def install_dropins() -> None:
    """Make `import pointnet2_cuda / iou3d_cuda / roipool3d_cuda` resolve to the B200 modules."""
    for name in ("pointnet2_cuda", "iou3d_cuda", "roipool3d_cuda"):
        sys.modules[name] = importlib.import_module(f"ws3d_b200.dropin.{name}")


def library_path() -> str:
    from . import _C
    return _C.LIB_PATH
