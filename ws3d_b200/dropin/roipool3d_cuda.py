"""Drop-in for the reference extension module `roipool3d_cuda`
(lib/utils/roipool3d/src/roipool3d.cpp:199-202).  `forward_slow` is the same computation as
`forward` in the reference (one-kernel vs three-kernel path); both map to the single B200 kernel."""
from ws3d_b200.native import pts_in_boxes3d_cpu, roipool3d_cpu  # noqa: F401
from ws3d_b200.native import roipool3d_forward as forward  # noqa: F401
from ws3d_b200.native import roipool3d_forward as forward_slow  # noqa: F401
