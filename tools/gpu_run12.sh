set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_v18.json 2> gpurun_out/bench_v18.err; tail -c 300 gpurun_out/bench_v18.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v18.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['rpn'], d['single_batch_latency'], d['two_in_flight'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])
print(d['roofline'])
for k in d['kernels'][:12]: print(k)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_v18.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches_v18.csv
