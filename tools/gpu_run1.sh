set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r1s3.log; tail -5 gpurun_out/pytest_gpu_r1s3.log
timeout 600 python bench.py > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; tail -c 600 gpurun_out/bench_v11.err; head -c 1500 gpurun_out/bench_v11.json
for sb in 84 100 116 132; do timeout 300 python bench.py --no-cpu-baseline --sm-budget $sb > gpurun_out/bench_v11_sb$sb.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_v11_sb$sb.json'));print($sb, d['value'], d['ms_per_step'], d['e2e']['value'], d['single_batch_latency'])"; done
timeout 600 python tools/fps_sweep.py 2>&1 | tail -20
timeout 300 python tools/stage2_bench.py > gpurun_out/stage2_v2.json 2>gpurun_out/stage2.err; cat gpurun_out/stage2_v2.json
timeout 300 python tools/train_rpn_bench.py > gpurun_out/train_rpn_v2.json 2>gpurun_out/train.err; cat gpurun_out/train_rpn_v2.json
