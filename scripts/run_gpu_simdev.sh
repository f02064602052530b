#!/bin/bash
# gpurun wrapper (development): integer-simulation tests, simulation bench, per-layer conv tables.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_intsim.py tests/test_gpu_int8_pipeline.py -m gpu -q > gpurun_out/pytest_sim.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sim.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/pytest_sim.log | tail -15
timeout 900 python bench_sim.py --mode ${MODE:-model} > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "sim rc=$?"
tail -3 gpurun_out/bench_sim.err | cut -c1-300; cat gpurun_out/bench_sim.json
timeout 600 python bench_conv_layers.py --s8-out > gpurun_out/conv_layers_s8.txt 2>&1
cat gpurun_out/conv_layers_s8.txt
timeout 600 python bench_conv_layers.py > gpurun_out/conv_layers.txt 2>&1
tail -2 gpurun_out/conv_layers.txt
