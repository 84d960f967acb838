mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 300 ncu --set full --clock-control none --import-source on -k regex:three_interpolate_smem4 -s 3 -c 1 -o /tmp/interp python tools/interp_bench.py > gpurun_out/ncu_interp.log 2>&1; tail -2 gpurun_out/ncu_interp.log | cut -c1-200
ncu -i /tmp/interp.ncu-rep --page source --csv > gpurun_out/interp_src.csv 2>/dev/null
ncu -i /tmp/interp.ncu-rep --page raw --csv > gpurun_out/interp_raw.csv 2>/dev/null
ls -la gpurun_out/interp_*.csv
