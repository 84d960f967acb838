set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 300 python -m pytest tests/test_gpu_modules.py tests/test_gpu_mlp.py -x -q 2>&1 | tail -3
for cfg in "1 1" "1 2" "1 3" "2 1" "2 2"; do set -- $cfg; WS3D_SCALE_STREAMS=$1 timeout 300 python bench.py --no-cpu-baseline --feature-streams $2 > gpurun_out/bench_v20.json 2>gpurun_out/bench_v20.err; tail -c 300 gpurun_out/bench_v20.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v20.json'));print('scale_streams $1 feature_streams $2:', d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn']['scenes_per_s'], d['two_in_flight']['ms_per_step'], d['single_batch_latency']['ms'])"; done
