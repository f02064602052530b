#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_int8_pipeline.py -m gpu -x -q > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pipe.log
tail -30 gpurun_out/pytest_pipe.log
