#!/usr/bin/env bash
# sampler variant x pipeline knobs: steady-state Mpoints/s of the streamed backbone
run() {  # smem clouds_per_cta_cap inflight fstreams budget cold_start [shape]
  WS3D_FPS_SMEM=$1 WS3D_FPS_SMEM_CLOUDS=$2 WS3D_FPS_SMEM_SHAPE=${7:-0} WS3D_FPS_PAIR=${8:-1} python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --inflight $3 --feature-streams $4 --sm-budget $5 --cold-start $6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('smem $1 clouds $2 inflight $3 fstreams $4 budget $5 cold $6 shape ${7:-0} pair ${8:-1}:', d['value'], d['steady_state']['Mpoints_per_s'], d['rpn']['steady_state_scenes_per_s'], d['verify']['streamed_checksums_equal_plain_forward'])
"
}
for cfg in "${@}"; do run $cfg; done
