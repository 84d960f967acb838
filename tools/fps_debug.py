import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from ws3d_b200 import native
dev = "cuda:0"
def cloud(rng, b, n, kind):
    if kind == "uniform": return rng.uniform(-10, 10, (b, n, 3)).astype(np.float32)
    if kind == "grid": return rng.integers(-3, 4, (b, n, 3)).astype(np.float32)
for (b, n, m, kind) in [(2, 2048, 128, "uniform"), (2, 5000, 700, "grid"), (2, 2048, 512, "grid")]:
    rng = np.random.default_rng(b * 7919 + n + m)
    xyz = cloud(rng, b, n, kind)
    exp_idx, exp_temp = oracle.furthest_point_sample(xyz, m, return_temp=True)
    x = torch.from_numpy(xyz).to(dev)
    temp = torch.full((b, n), 1e10, device=dev); idx = torch.empty((b, m), dtype=torch.int32, device=dev); nx = torch.empty((b, m, 3), device=dev)
    native.set_fps_mode(1)
    native.furthest_point_sampling_gather(b, n, m, x, temp, idx, nx)
    native.set_fps_mode(0)
    gi, gt = idx.cpu().numpy(), temp.cpu().numpy()
    bad = np.argwhere(gi != exp_idx)
    print(kind, n, m, "idx mismatches", len(bad), "first", bad[:3].tolist(), "temp mismatches", int((gt != exp_temp).sum()))
    if len(bad):
        c, k = bad[0]
        print("  at sample", k, "got", gi[c, k], xyz[c, gi[c, k]], "want", exp_idx[c, k], xyz[c, exp_idx[c, k]], "prev", gi[c, k-1])
    tb = np.argwhere(gt != exp_temp)
    for c, k in tb[:3]:
        print("  temp at", c, k, xyz[c, k], "got", gt[c, k], "want", exp_temp[c, k])
