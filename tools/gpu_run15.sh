set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 600 python -m pytest tests/test_gpu_pointnet2.py -x -q -k "fps" 2>&1 | tail -5
timeout 600 python tools/fps_mode_bench.py 2>&1 | tail -6
