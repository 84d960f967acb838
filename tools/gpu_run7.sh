set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q -k "fused_sa" 2>&1 | tail -5
timeout 300 python tools/sa_fused_bench.py 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_mlp_fused -c 4 -o gpurun_out/sa_fused_v3 python tools/sa_fused_bench.py --one > gpurun_out/ncu_sa.log 2>&1; tail -3 gpurun_out/ncu_sa.log
