#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_intsim.py -m gpu -x -q > gpurun_out/pytest_intsim.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_intsim.log
tail -40 gpurun_out/pytest_intsim.log
