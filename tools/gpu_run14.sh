set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
date
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench_v18.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches_bench_v18.csv
date
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_two_phase_v18.csv python tools/prof_two_phase.py 3 > gpurun_out/ncu_tp.log 2>&1; wc -l gpurun_out/launches_two_phase_v18.csv
date
timeout 600 ncu --set full --clock-control none -k regex:"fps_bucket|fps_flat|sa_mlp_fused|mlp_layer_kernel|three_interpolate|group_concat|three_nn_grid|grid_query" -s 40 -c 60 -o /tmp/r1v18_full python tools/prof_two_phase.py 2 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-200
ls -la /tmp/r1v18_full.ncu-rep
ncu -i /tmp/r1v18_full.ncu-rep --page raw --csv > gpurun_out/r1v18_full_raw.csv 2>/dev/null; ls -la gpurun_out/
gzip -f gpurun_out/launches_bench_v18.csv
date
