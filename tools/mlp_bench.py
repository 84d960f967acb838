"""Per-layer timings of ws3d_mlp_layer at the backbone's shapes (B=16) for each tile configuration."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYERS = [  # (c1, c2, c_out, cols, pool)
    (4, 0, 32, 131072, 0), (32, 0, 32, 131072, 0), (32, 0, 64, 131072, 32),
    (4, 0, 16, 65536, 0), (16, 0, 16, 65536, 0), (16, 0, 32, 65536, 16),
    (99, 0, 64, 32768, 0), (64, 0, 96, 32768, 0), (96, 0, 128, 32768, 32),
    (99, 0, 64, 16384, 0), (64, 0, 64, 16384, 0), (64, 0, 128, 16384, 16),
    (259, 0, 128, 8192, 0), (128, 0, 196, 8192, 0), (196, 0, 256, 8192, 32),
    (259, 0, 128, 4096, 0), (128, 0, 196, 4096, 0), (196, 0, 256, 4096, 16),
    (515, 0, 256, 2048, 0), (256, 0, 384, 2048, 0), (384, 0, 512, 2048, 32),
    (515, 0, 256, 1024, 0), (256, 0, 256, 1024, 0), (256, 0, 512, 1024, 16),
    (1024, 512, 512, 256, 0), (512, 0, 512, 256, 0), (512, 256, 512, 1024, 0), (512, 0, 512, 1024, 0),
    (512, 96, 256, 4096, 0), (256, 0, 256, 4096, 0), (256, 1, 128, 16384, 0), (128, 0, 128, 16384, 0),
]
CHILD = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from ws3d_b200 import native
LAYERS = %r
dev = "cuda:0"; B = 16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for (c1, c2, co, cols, pool) in LAYERS:
    k1 = (c1 + 31) // 32 * 32; k2 = (c2 + 31) // 32 * 32 if c2 else 0
    cop = (co + 127) // 128 * 128
    w = torch.randn(cop, k1 + k2, device=dev); sh = torch.randn(cop, device=dev)
    x1 = torch.randn(B, c1, cols, device=dev); x2 = torch.randn(B, c2, cols, device=dev) if c2 else None
    y = torch.empty(B, co, cols // pool if pool else cols, device=dev)
    rep = 2 if co <= 32 else (1 if co <= 64 else 0)
    def run(): native.mlp_layer(B, co, cop, c1, c2, cols, w, sh, x1, x2, y, 3 | (rep << 4), pool)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    out.append(float(np.median(ts)))
print(json.dumps(out))
''' % (ROOT, LAYERS)
res = {}
for cfg in sys.argv[1:] or ["-1", "0", "3"]:
    e = dict(os.environ); e["WS3D_MLP_CFG"] = cfg
    p = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True)
    if p.returncode: print("cfg", cfg, "FAILED", p.stderr[-1500:]); continue
    res[cfg] = json.loads(p.stdout.strip().splitlines()[-1])
cfgs = sorted(res)
print("layer (c1,c2,co,cols,pool)".ljust(36) + "".join(f"cfg{c:>2s} ms  GB/s   " for c in cfgs))
tot = {c: 0.0 for c in cfgs}
for i, L in enumerate(LAYERS):
    c1, c2, co, cols, pool = L
    byt = 4 * 16 * ((c1 + c2) * cols + co * (cols // pool if pool else cols))
    line = str(L).ljust(36)
    for c in cfgs:
        ms = res[c][i]; tot[c] += ms
        line += f"{ms:7.4f} {byt / ms / 1e6:6.0f}   "
    print(line)
print("total".ljust(36) + "".join(f"{tot[c]:7.3f}          " for c in cfgs))
json.dump({"layers": LAYERS, "ms": res}, open(os.path.join(ROOT, "gpurun_out", "mlp_bench.json"), "w"))
