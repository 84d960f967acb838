set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for sb in 84 0 116; do timeout 300 python bench.py --no-cpu-baseline --sm-budget $sb > gpurun_out/bench_v14_sb$sb.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_v14_sb$sb.json'));print($sb, d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn']['scenes_per_s'], d['single_batch_latency'])"; done
timeout 300 python tools/stage2_bench.py > gpurun_out/stage2_v3.json 2>gpurun_out/stage2.err; cat gpurun_out/stage2_v3.json
