#!/bin/bash
# gpurun wrapper (round 2, experiment): fused conv + add epilogue after hoisting the per-slab divisions
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_int8_pipeline.py tests/test_gpu_c4_at_size.py -m gpu -x -q > gpurun_out/pytest_fa.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_fa.log | cut -c1-200
for r in 3 7 13 19; do timeout 200 python bench_conv_layers.py --s8-out --fused-add --only $r 2>&1 | grep "^(" ; done | tee gpurun_out/fused_layers_exp5.txt
timeout 600 python bench_sim.py --mode model > gpurun_out/bench_sim_exp5.json 2> gpurun_out/bench_sim_exp5.err; echo "sim rc=$?"
cut -c1-1500 gpurun_out/bench_sim_exp5.json
