mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_modules.py -x -q 2>&1 | tail -8
for pm in 1 0; do WS3D_SA_PREMUL=$pm timeout 300 python bench.py --no-cpu-baseline | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('sa premul $pm:', d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn']['scenes_per_s'], d['single_batch_latency']['ms'])"; done
