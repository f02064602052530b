#!/bin/bash
# gpurun wrapper: the whole GPU suite (what the driver runs at round end), log under gpurun_out/.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=15 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | tail -30
