mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_v23.json 2> gpurun_out/bench_v23.err; tail -c 300 gpurun_out/bench_v23.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v23.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn'], d['single_batch_latency']['ms'], d['two_in_flight']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])
PY
WS3D_SCALE_STREAMS=0 timeout 300 python bench.py --no-cpu-baseline | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('no scale/head streams:', d['value'], d['rpn']['scenes_per_s'])"
