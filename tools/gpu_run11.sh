set -x
mkdir -p gpurun_out
for inf in 4 5 6 7; do timeout 300 python bench.py --no-cpu-baseline --inflight $inf > gpurun_out/bench_v17_in$inf.json 2>gpurun_out/bench_v17.err; tail -c 300 gpurun_out/bench_v17.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v17_in$inf.json'));print($inf, d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn']['scenes_per_s'], d['single_batch_latency']['ms'])"; done
