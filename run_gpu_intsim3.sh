#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_intsim.py tests/test_gpu_parity.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench_conv_layers.py > gpurun_out/conv_layers.txt 2>&1; echo rc=$?
head -3 gpurun_out/conv_layers.txt; tail -2 gpurun_out/conv_layers.txt
timeout 1200 python bench_sim.py --mode both > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "rc=$?"
cat gpurun_out/bench_sim.json
