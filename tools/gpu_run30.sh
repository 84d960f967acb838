mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
