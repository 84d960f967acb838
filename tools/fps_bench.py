"""FPS timings on the GPU box: default dispatch, register/cluster kernel (fps.cu), bucketed single-CTA kernel (fps_bucket.cu).
Each variant runs in a child process because the kernel choice is read from the environment once."""
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
out = []
for (b, n, m) in [(16, 16384, 4096), (1, 16384, 4096), (16, 4096, 1024), (16, 1024, 256), (32, 16384, 4096), (64, 16384, 4096), (96, 16384, 4096), (96, 4096, 1024)]:
    pts = torch.from_numpy(np.ascontiguousarray(synth.make_batch(min(b, 16), n)[..., :3])).to(dev)
    if b > 16: pts = pts.repeat(b // 16, 1, 1).contiguous()
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
    out.append({"b": b, "n": n, "m": m, "ms": float(np.median(ts)), "us_per_iter": float(np.median(ts)) * 1e3 / (m - 1),
                "checksum": int(idx.long().sum().item())})
print(json.dumps(out))
''' % ROOT


def main():
    res = {}
    for name, env in [("auto", {}), ("flat", {"WS3D_FPS_BUCKET": "0", "WS3D_FPS_FLAT": "1"}),
                      ("cluster", {"WS3D_FPS_BUCKET": "0", "WS3D_FPS_FLAT": "0"}), ("bucket", {"WS3D_FPS_BUCKET": "1", "WS3D_FPS_SMEM": "0"}),
                      ("smem", {"WS3D_FPS_BUCKET": "1", "WS3D_FPS_SMEM": "1"}),
                      ("smem_single", {"WS3D_FPS_BUCKET": "1", "WS3D_FPS_SMEM": "1", "WS3D_FPS_PAIR": "0"}),
                      ("smem8x2", {"WS3D_FPS_BUCKET": "1", "WS3D_FPS_SMEM": "1", "WS3D_FPS_PAIR": "0", "WS3D_FPS_SMEM_SHAPE": "1"})]:
        e = dict(os.environ); e.update(env)
        p = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True)
        if p.returncode:
            print(name, "FAILED", p.stderr[-2000:])
            continue
        res[name] = json.loads(p.stdout.strip().splitlines()[-1])
    names = [k for k in ("auto", "flat", "cluster", "bucket", "smem", "smem_single", "smem8x2") if k in res]
    for i in range(len(res[names[0]])):
        a = res[names[0]][i]
        line = f"b={a['b']:3d} n={a['n']:6d} m={a['m']:5d} "
        for k in names:
            r = res[k][i]
            line += f" | {k} {r['ms']:.3f}"
        line += "  same_idx=" + str(len({res[k][i]['checksum'] for k in names}) == 1)
        print(line)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "fps_bench.json"), "w"))


if __name__ == "__main__":
    main()
