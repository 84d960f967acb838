set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/bench_2gpu_v22.json 2> gpurun_out/bench_2gpu.err; tail -c 600 gpurun_out/bench_2gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_2gpu_v22.json').read().strip().splitlines()[-1]);print('2 GPUs:', d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn']['scenes_per_s'], d['n_gpus'])"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu_v22.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_1gpu_v22.json'));print('1 GPU:', d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn']['scenes_per_s'])"
