mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 600 python bench.py > gpurun_out/bench_v25.json 2> gpurun_out/bench_v25.err; tail -c 300 gpurun_out/bench_v25.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v25.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn'], d['single_batch_latency'], d['two_in_flight'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])
print(d['roofline'])
for k in d['kernels'][:16]: print(k)
PY
timeout 300 python tools/stage2_bench.py > gpurun_out/stage2_v4.json 2>/dev/null; cat gpurun_out/stage2_v4.json | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_two_phase_v25.csv python tools/prof_two_phase.py 3 > gpurun_out/ncu_tp.log 2>&1; wc -l gpurun_out/launches_two_phase_v25.csv
timeout 500 ncu --set full --clock-control none -k regex:"fps_bucket|sa_mlp_fused|mlp_layer_kernel|three_interpolate|group_affine|group_concat|three_nn_grid|grid_query" -s 36 -c 52 -o /tmp/r1v25_full python tools/prof_two_phase.py 2 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
ncu -i /tmp/r1v25_full.ncu-rep --page raw --csv > gpurun_out/r1v25_full_raw.csv 2>/dev/null; ls -la gpurun_out/ | head -20
