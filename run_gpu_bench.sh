#!/bin/bash
# gpurun wrapper: GPU tests, then the 1-GPU bench line (+ reference arm), then the microbenchmark.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${STEPS:-16} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
if [ -x pytorch-quantity_b200/csrc/bench/hist_microbench ]; then
  timeout 300 pytorch-quantity_b200/csrc/bench/hist_microbench > gpurun_out/hist_microbench.log 2>&1
fi
