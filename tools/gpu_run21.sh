mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 300 python tools/train_profile.py 2>&1 | tail -45 | cut -c1-220
