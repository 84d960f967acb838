mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 4 > gpurun_out/bench_4gpu_v25.json 2> gpurun_out/bench_4gpu.err; tail -c 400 gpurun_out/bench_4gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_4gpu_v25.json').read().strip().splitlines()[-1]);print('4 GPUs:', d['value'], d['ms_per_step'], d['e2e']['value'], d['rpn']['scenes_per_s'], d['n_gpus'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 | cut -c1-200
