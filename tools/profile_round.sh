#!/bin/bash
# The GPU-box recipe behind profiles/ (run under gpurun; outputs land in gpurun_out/, summarised here with tools/ncu_summary.py):
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/profile_round.sh'
#   python tools/ncu_summary.py launches gpurun_out/launches_two_phase.csv profiles/<name>_summary.csv --passes 3 --cmd "..."
#   python tools/ncu_summary.py full gpurun_out/full_raw.csv profiles/<name>.csv
# bench.py itself cannot run under ncu (graph capture of cluster launches fails in the profiler), so the launch list is
# taken from tools/prof_two_phase.py, which issues eagerly exactly what the streamed pipeline replays.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/stage2_bench.py > gpurun_out/stage2.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_two_phase.csv \
    python tools/prof_two_phase.py 3 > gpurun_out/ncu_tp.log 2>&1
# the report with sources is too large to travel back (64 MiB limit): export the raw page on the box
timeout 500 ncu --set full --clock-control none \
    -k regex:"fps_bucket|sa_mlp_fused|mlp_layer_kernel|three_interpolate|group_affine|group_concat|three_nn_grid|grid_query" \
    -s 36 -c 52 -o /tmp/full python tools/prof_two_phase.py 2 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/full.ncu-rep --page raw --csv > gpurun_out/full_raw.csv 2>/dev/null
