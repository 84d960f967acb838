set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_next_rows.py -x -q 2>&1 | tail -5
timeout 600 python tools/next_rows_bench.py 2>&1 | grep cylinder
timeout 600 python bench.py --sm-budget 84 > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err; tail -c 400 gpurun_out/bench_v12.err
python -c "
import json;d=json.load(open('gpurun_out/bench_v12.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn'], d['cpu_baseline'])"
for fm in 8192; do WS3D_FPS_FLAT_MIN=$fm timeout 300 python bench.py --no-cpu-baseline --sm-budget 84 > gpurun_out/bench_v12_fps2single.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_v12_fps2single.json'));print('fps2 single cta', d['value'], d['ms_per_step'], d['e2e']['value'], d['single_batch_latency'])"; done
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 | head -c 1200
