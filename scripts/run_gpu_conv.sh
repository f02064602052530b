#!/bin/bash
# gpurun wrapper: integer-simulation tests, per-layer conv table, simulation bench, ncu of selected conv layers.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_intsim.py tests/test_gpu_int8_pipeline.py -x -q > gpurun_out/pytest_conv.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_conv.log
timeout 600 python bench_conv_layers.py --s8-out > gpurun_out/conv_layers_s8.txt 2>&1
cat gpurun_out/conv_layers_s8.txt
timeout 600 python bench_conv_layers.py > gpurun_out/conv_layers.txt 2>&1
tail -2 gpurun_out/conv_layers.txt
timeout 900 python bench_sim.py --mode model > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "sim rc=$?"
tail -3 gpurun_out/bench_sim.err | cut -c1-300; cat gpurun_out/bench_sim.json
for L in ${NCU_LAYERS:-0 2 3}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_s8_kernel|conv_rows_s8_kernel" -s 3 -c 1 \
      -f -o gpurun_out/prof_conv_layer$L python bench_conv_layers.py --s8-out --only $L > gpurun_out/ncu_layer$L.log 2>&1
  echo "layer $L rc=$?"
done
