#!/bin/bash
# gpurun wrapper: live GPU-vs-GPU parity against the unmodified reference (baseline/_ref) + the rest of the GPU suite.
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 1700 python -m pytest tests/test_gpu_vs_reference.py -m gpu -q --durations=25 ${PYTEST_ARGS} > gpurun_out/pytest_vs_ref.log 2>&1; echo "vs_ref rc=$?" >> gpurun_out/pytest_vs_ref.log
tail -40 gpurun_out/pytest_vs_ref.log
