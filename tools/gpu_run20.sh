set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 300 python tools/train_rpn_bench.py --graph 0 2>&1 | tail -2 | cut -c1-700
timeout 300 python tools/train_rpn_bench.py --graph 1 2>&1 | tail -4 | cut -c1-900
