"""Eager passes of what the streamed pipeline runs, at the bench shapes (B=16 x 16384 points) -- the command ncu wraps:
coordinate phase (FPS in throughput mode, ball queries, stencils) then feature phase (fused SA scales, MLP layers,
interpolation).  `python tools/prof_two_phase.py [passes]`; one warm-up pass precedes cudaProfilerStart, so with
`ncu --profile-from-start off` the launch list holds steady-state passes only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ws3d_b200 import models, native, synth

torch.manual_seed(0)
dev = "cuda:0"
model = models.Pointnet2MSG(input_channels=1).to(dev).eval()
pts = torch.from_numpy(synth.make_batch(16, 16384)).to(dev)


def one_pass():
    native.set_fps_mode(1)
    plan = model.coordinate_phase(pts)
    native.set_fps_mode(0)
    return model.feature_phase(pts, plan)[1]


with torch.no_grad():
    one_pass()                         # one-time work (BatchNorm folded into the weights, scratch growth): outside the capture
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()          # ncu --profile-from-start off captures from here
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        out = one_pass()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok", float(out.abs().mean()))
