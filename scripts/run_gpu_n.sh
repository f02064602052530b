#!/bin/bash
# gpurun wrapper: headline bench on N GPUs of one box (N from $NGPU), as the driver launches it
mkdir -p gpurun_out
N=${NGPU:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps ${STEPS:-16} --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; tail -3 gpurun_out/bench_${N}gpu.err | cut -c1-300; cat gpurun_out/bench_${N}gpu.json
