set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for ss in 1 0; do WS3D_SCALE_STREAMS=$ss timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v19_ss$ss.json 2>gpurun_out/bench_v19.err; tail -c 300 gpurun_out/bench_v19.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v19_ss$ss.json'));print('scale_streams', $ss, d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn']['scenes_per_s'], d['two_in_flight']['ms_per_step'], d['single_batch_latency']['ms'])"; done
