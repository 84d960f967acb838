#!/usr/bin/env bash
# pipeline knob sweep: steady-state Mpoints/s of the streamed backbone
for cfg in "6 2 100" "6 2 116" "6 2 132" "6 2 148" "6 2 84" "7 2 100" "8 2 100" "6 3 100" "7 3 116" "5 2 100" "6 1 100" "8 3 132"; do
  set -- $cfg
  python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --inflight $1 --feature-streams $2 --sm-budget $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('inflight $1 fstreams $2 budget $3 :', d['value'], d['steady_state']['Mpoints_per_s'], d['rpn']['steady_state_scenes_per_s'])
"
done
