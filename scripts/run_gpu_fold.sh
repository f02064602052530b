#!/bin/bash
# gpurun wrapper: folded-bias epilogue -- tests, per-layer table with and without the fold, simulation bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_intsim.py tests/test_gpu_int8_pipeline.py tests/test_gpu_lenet.py -x -q > gpurun_out/pytest_conv.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_conv.log
timeout 600 python bench_conv_layers.py --s8-out > gpurun_out/conv_layers_s8.txt 2>&1
cat gpurun_out/conv_layers_s8.txt
timeout 600 python bench_conv_layers.py --s8-out --classic-bias > gpurun_out/conv_layers_s8_classic.txt 2>&1
cat gpurun_out/conv_layers_s8_classic.txt
timeout 600 python bench_conv_layers.py > gpurun_out/conv_layers.txt 2>&1
tail -2 gpurun_out/conv_layers.txt
timeout 900 python bench_sim.py --mode model > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "sim rc=$?"
cat gpurun_out/bench_sim.json | cut -c1-330
