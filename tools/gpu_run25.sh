mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_pointnet2.py -x -q 2>&1 | tail -3
timeout 300 python tools/train_rpn_bench.py > gpurun_out/train_rpn_v3.json 2>/dev/null; cat gpurun_out/train_rpn_v3.json | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/train_rpn_bench.py > gpurun_out/train_rpn_2gpu_v3.json 2>gpurun_out/train2.err; tail -1 gpurun_out/train_rpn_2gpu_v3.json | cut -c1-400
