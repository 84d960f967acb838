"""Active buckets and update rounds per iteration of the shared-memory throughput sampler (csrc/fps_smem.cu, debug build of the
kernel: WS3D_FPS_STATS=1).  `python tools/fps_stats.py`"""
import os
import sys

os.environ["WS3D_FPS_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ws3d_b200 import native, pointnet2_utils, synth

pts = torch.from_numpy(np.ascontiguousarray(synth.make_batch(16, 16384)[..., :3])).to("cuda:0")
native.set_fps_mode(1)
for m in (4096, 1024, 256):
    pointnet2_utils.sample_and_gather(pts, m)
torch.cuda.synchronize()
