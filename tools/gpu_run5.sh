set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err; tail -c 400 gpurun_out/bench_v13.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v13.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn'], d['single_batch_latency'])
for k in d['kernels'][:14]: print(k)
print([k for k in d['kernels'] if k['kernel']=='sa_mlp_fused'])
PY
WS3D_SA_FUSED=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v13_unfused.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/bench_v13_unfused.json'));print('unfused', d['value'], d['ms_per_step'], d['single_batch_latency'])"
for sb in 0 116; do timeout 300 python bench.py --no-cpu-baseline --sm-budget $sb > gpurun_out/bench_v13_sb$sb.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_v13_sb$sb.json'));print($sb, d['value'], d['ms_per_step'], d['e2e']['value'])"; done
timeout 300 python tools/stage2_bench.py > gpurun_out/stage2_v3.json 2>gpurun_out/stage2.err; cat gpurun_out/stage2_v3.json
