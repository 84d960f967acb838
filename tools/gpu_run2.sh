set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_next_rows.py tests/test_gpu_iou3d.py -x -q 2>&1 | tail -15
timeout 600 python tools/next_rows_bench.py 2>&1 | tail -12
