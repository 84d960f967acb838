"""Eager passes of what the streamed pipeline runs, at the bench shapes (B=16 x 16384 points) -- the command ncu wraps:
coordinate phase (FPS in throughput mode, ball queries, stencils) then feature phase (fused SA scales, MLP layers,
interpolation).  `python tools/prof_two_phase.py [passes]`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ws3d_b200 import models, native, synth

torch.manual_seed(0)
dev = "cuda:0"
model = models.Pointnet2MSG(input_channels=1).to(dev).eval()
pts = torch.from_numpy(synth.make_batch(16, 16384)).to(dev)
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        native.set_fps_mode(1)
        plan = model.coordinate_phase(pts)
        native.set_fps_mode(0)
        out = model.feature_phase(pts, plan)[1]
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
