set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q -k "fused_sa" 2>&1 | tail -3
timeout 300 python tools/sa_fused_bench.py 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v15.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_v15.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn']['scenes_per_s'], d['single_batch_latency'])"
