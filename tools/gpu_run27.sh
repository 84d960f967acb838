mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 300 python -m pytest tests/test_gpu_next_rows.py -x -q 2>&1 | tail -3
timeout 300 python tools/next_rows_bench.py 2>&1 | tail -2
