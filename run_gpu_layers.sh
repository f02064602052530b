#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench_conv_layers.py > gpurun_out/conv_layers.txt 2>&1; echo rc=$?
timeout 900 python bench_conv_layers.py --s8-out > gpurun_out/conv_layers_s8.txt 2>&1; echo rc=$?
cat gpurun_out/conv_layers.txt; tail -3 gpurun_out/conv_layers_s8.txt
