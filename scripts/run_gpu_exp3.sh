#!/bin/bash
# gpurun wrapper (round 2, experiment): patch-window 3x3 path -- parity tests, then the two ResNet-50 layers it serves,
# with and without it.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_intsim.py -m gpu -x -q -k "patch_windows" > gpurun_out/pytest_win.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_win.log | cut -c1-200
for r in 2 10; do
  timeout 200 python bench_conv_layers.py --s8-out --only $r 2>&1 | grep "^("
  timeout 200 python bench_conv_layers.py --s8-out --only $r --no-windows 2>&1 | grep "^("
done | tee gpurun_out/win_layers.txt
