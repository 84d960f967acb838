import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from ws3d_b200 import native, synth
dev = "cuda:0"
b, n, m = 16, 16384, 4096
pts = torch.from_numpy(np.ascontiguousarray(synth.make_batch(b, n)[..., :3])).to(dev)
idx = torch.empty((b, m), dtype=torch.int32, device=dev)
new_xyz = torch.empty((b, m, 3), device=dev)
for _ in range(2):
    native.furthest_point_sampling_gather(b, n, m, pts, None, idx, new_xyz)
torch.cuda.synchronize()
