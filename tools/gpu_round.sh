#!/usr/bin/env bash
# One GPU-box visit: parity tests, the bench line, sanitizer evidence, the ncu launch list.  Everything lands in gpurun_out/.
# usage: tools/gpu_round.sh <tag> [steps...]   (steps: tests bench sanitize ncu; default all)
set -u
TAG=${1:-r2}; shift || true
STEPS=${*:-tests bench sanitize ncu}
mkdir -p gpurun_out
for s in $STEPS; do
  case $s in
    tests)    timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log ;;
    bench)    timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench.err; head -c 1500 gpurun_out/${TAG}_bench.json ;;
    sanitize) timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize.py --quick > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_memcheck.log
              timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize.py --quick > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_racecheck.log ;;
    ncu)      timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,sm__cycles_active.avg,launch__grid_size,launch__block_size,launch__occupancy_limit_blocks --clock-control none --profile-from-start off -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_two_phase.py 3 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log ;;
    ncufull)  timeout 900 ncu --set full --clock-control none --profile-from-start off --import-source on -k regex:"sa_mlp_fused|mlp_layer|fps_bucket|roipool3d|three_interpolate|group_affine" -c 40 -o gpurun_out/${TAG}_full python tools/prof_two_phase.py 1 > gpurun_out/${TAG}_ncufull.log 2>&1; echo "ncufull rc=$?" ;;
  esac
done
