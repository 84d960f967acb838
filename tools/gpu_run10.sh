set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_modules.py -x -q 2>&1 | tail -3
for inf in 3 4; do timeout 300 python bench.py --no-cpu-baseline --inflight $inf > gpurun_out/bench_v16_in$inf.json 2>gpurun_out/bench_v16.err; tail -c 300 gpurun_out/bench_v16.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v16_in$inf.json'));print($inf, d['value'], d['ms_per_step'], d['e2e']['value'], d['two_in_flight'], d['single_batch_latency']['ms'])"; done
