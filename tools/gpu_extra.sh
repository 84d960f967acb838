#!/usr/bin/env bash
# Second half of a GPU-box visit: the measurements that are not part of bench.py.  usage: tools/gpu_extra.sh <tag> [steps...]
set -u
TAG=${1:-r2}; shift || true
STEPS=${*:-loader trainprof opsfull}
mkdir -p gpurun_out
for s in $STEPS; do
  case $s in
    loader)    timeout 600 python tools/loader_harness.py > gpurun_out/${TAG}_loader.json 2> gpurun_out/${TAG}_loader.err; echo "loader rc=$?"; tail -c 400 gpurun_out/${TAG}_loader.err; head -c 1200 gpurun_out/${TAG}_loader.json ;;
    trainprof) timeout 600 python tools/train_profile.py 32 > gpurun_out/${TAG}_train_profile.txt 2>&1; echo "trainprof rc=$?"; head -30 gpurun_out/${TAG}_train_profile.txt ;;
    opsfull)   timeout 900 ncu --set full --clock-control none --import-source on -c 60 -o gpurun_out/${TAG}_ops_full python tools/prof_ops.py > gpurun_out/${TAG}_opsfull.log 2>&1; echo "opsfull rc=$?"
               ncu -i gpurun_out/${TAG}_ops_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ops_full_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_ops_full.ncu-rep ;;
    train1)    timeout 600 python tools/train_rpn_bench.py --batch 32 --steps 8 > gpurun_out/${TAG}_train_1gpu.json 2> gpurun_out/${TAG}_train_1gpu.err; echo "train1 rc=$?"; cat gpurun_out/${TAG}_train_1gpu.json ;;
  esac
done
