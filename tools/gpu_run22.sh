mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
WS3D_NATIVE_BN=0 timeout 300 python tools/train_rpn_bench.py --graph 0 2>&1 | tail -1 | cut -c1-250
timeout 300 python tools/train_rpn_bench.py --graph 0 2>&1 | tail -1 | cut -c1-250
timeout 300 python tools/train_rpn_bench.py --graph 1 2>&1 | tail -1 | cut -c1-250
timeout 300 python tools/train_profile.py 2>&1 | tail -34 | cut -c1-200 | awk '{print $1,$2,$3, $(NF-3), $(NF-2)}' | head -24
