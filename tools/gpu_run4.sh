set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q -k "fused_sa" 2>&1 | tail -30
