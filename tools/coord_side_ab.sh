for side in 1 0 1 0; do
WS3D_COORD_SIDE=$side python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('coord_side $side :', d['value'], d['e2e']['value'], d['steady_state']['Mpoints_per_s'], d['rpn']['scenes_per_s'], d['verify']['streamed_checksums_equal_plain_forward'])
"
done
