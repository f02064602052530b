#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_intsim.py tests/test_gpu_parity.py -m gpu -x -q -k "intsim or quantize or gemm or conv or linear or recon" > gpurun_out/pytest_intsim.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_intsim.log
tail -15 gpurun_out/pytest_intsim.log
timeout 600 python bench_conv_layers.py > gpurun_out/conv_layers.txt 2>&1; echo rc=$?
cat gpurun_out/conv_layers.txt
timeout 600 python bench_conv_layers.py --s8-out 2>&1 | tail -2
