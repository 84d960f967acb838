"""Decomposition sweep of the one-level FPS kernel (fps_flat_kernel) on the GPU box: clusters C x T threads x P points
for level 1 (16384 -> 4096) and level 2 (4096 -> 1024) at b = 16.  Each variant runs in a child process because the
kernel choice is read from the environment.  Writes gpurun_out/fps_sweep.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from ws3d_b200 import native, synth
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
b, n, m = [int(v) for v in sys.argv[1:4]]
pts = torch.from_numpy(np.ascontiguousarray(synth.make_batch(b, n)[..., :3])).to(dev)
idx = torch.empty((b, m), dtype=torch.int32, device=dev)
new_xyz = torch.empty((b, m, 3), device=dev)
def run():
    native.furthest_point_sampling_gather(b, n, m, pts, None, idx, new_xyz)
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(7):
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record(); e.synchronize()
    ts.append(s.elapsed_time(e))
print(json.dumps({"ms": float(np.median(ts)), "checksum": int(idx.long().sum().item())}))
''' % ROOT


def main():
    out = []
    for (b, n, m) in [(16, 16384, 4096), (16, 4096, 1024)]:
        ppts = (16, 8, 4) if n > 8192 else (8, 4, 2)
        for C in (2, 4, 8):
            for ppt in ppts:
                env = dict(os.environ, WS3D_FPS_BUCKET="0", WS3D_FPS_FLAT="1", WS3D_FPS_C=str(C), WS3D_FPS_PPT=str(ppt))
                p = subprocess.run([sys.executable, "-c", CHILD, str(b), str(n), str(m)], env=env, capture_output=True,
                                   text=True, timeout=300)
                T = (n // C) // ppt
                rec = {"b": b, "n": n, "m": m, "C": C, "T": T, "P": ppt}
                if p.returncode:
                    rec["error"] = p.stderr[-300:]
                else:
                    rec.update(json.loads(p.stdout.strip().splitlines()[-1]))
                    rec["us_per_iter"] = rec["ms"] * 1e3 / (m - 1)
                print(rec, flush=True)
                out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fps_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
