mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 300 python tools/interp_bench.py 2>&1 | tail -6
timeout 300 python -m pytest tests/test_gpu_pointnet2.py tests/test_gpu_mlp.py -x -q -k "interpolate or fp_module" 2>&1 | tail -2
