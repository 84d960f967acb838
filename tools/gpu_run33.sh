mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
python - <<'PY'
import os, numpy as np, torch
from ws3d_b200 import native, synth
dev="cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
native.set_fps_mode(1)
for (b,n,m) in [(16,16384,4096),(16,4096,1024),(16,8192,2048)]:
    pts = torch.from_numpy(np.ascontiguousarray(synth.make_batch(b, n)[..., :3])).to(dev)
    idx = torch.empty((b, m), dtype=torch.int32, device=dev); nx = torch.empty((b, m, 3), device=dev)
    run = lambda: native.furthest_point_sampling_gather(b, n, m, pts, None, idx, nx)
    for _ in range(3): run()
    torch.cuda.synchronize(); ts=[]
    for _ in range(5):
        flush.fill_(1); s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    print(b,n,m, round(float(np.median(ts)),3), int(idx.long().sum()))
PY
timeout 300 python -m pytest tests/test_gpu_pointnet2.py -x -q -k "fps" 2>&1 | tail -2
