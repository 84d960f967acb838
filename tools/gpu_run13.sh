set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_v18.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300; wc -l gpurun_out/launches_bench_v18.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_two_phase_v18.csv python tools/prof_two_phase.py 3 > gpurun_out/ncu_tp.log 2>&1; wc -l gpurun_out/launches_two_phase_v18.csv
timeout 900 ncu --set full --clock-control none --import-source on -s 75 -c 75 -o gpurun_out/r1v18_full python tools/prof_two_phase.py 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
