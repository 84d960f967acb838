set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 300 python -m pytest tests/test_gpu_modules.py -x -q 2>&1 | tail -3
for cfg in "5 2" "6 2" "7 2" "7 3"; do set -- $cfg; timeout 300 python bench.py --no-cpu-baseline --inflight $1 --feature-streams $2 > gpurun_out/bench_v21.json 2>gpurun_out/bench_v21.err; tail -c 300 gpurun_out/bench_v21.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v21.json'));print('inflight $1 feature_streams $2:', d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn']['scenes_per_s'])"; done
