#!/bin/bash
# gpurun wrapper (round 2): ncu captures of the plain int8 epilogue on the 1x1 64->256 expansion and of the stem kernel
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 2 -c 1 \
    -f -o gpurun_out/prof_gemm_s8_conv1x1_64to256x56 python bench_conv_layers.py --s8-out --only 3 > gpurun_out/ncu_a.log 2>&1
echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_rows_s8_kernel -s 2 -c 1 \
    -f -o gpurun_out/prof_conv_rows_stem python bench_conv_layers.py --s8-out --only 0 > gpurun_out/ncu_b.log 2>&1
echo "rc=$?"
