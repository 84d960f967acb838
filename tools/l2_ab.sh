for mode in flush inputs flush inputs; do
python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --l2 $mode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$mode', d['value'], d['e2e']['value'], d['steady_state']['Mpoints_per_s'], d['rpn']['steady_state_scenes_per_s'], d['verify']['streamed_checksums_equal_plain_forward'])
"
done
