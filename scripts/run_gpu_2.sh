#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "rc=$?"; tail -3 gpurun_out/bench_2gpu.err | cut -c1-300; cat gpurun_out/bench_2gpu.json
python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_k8.json 2>/dev/null; cat gpurun_out/bench_1gpu_k8.json
