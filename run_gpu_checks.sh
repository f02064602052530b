#!/bin/bash
# Convenience wrapper for gpurun: GPU tests, smoke and the statistics microbenchmark.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
if [ -x pytorch-quantity_b200/csrc/bench/hist_microbench ]; then
  timeout 300 pytorch-quantity_b200/csrc/bench/hist_microbench > gpurun_out/hist_microbench.log 2>&1
fi
tail -5 gpurun_out/smoke.log; tail -25 gpurun_out/pytest_gpu.log
