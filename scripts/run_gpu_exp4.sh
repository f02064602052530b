#!/bin/bash
# gpurun wrapper (round 2): ncu capture of the patch-window kernel on the 3x3 C=64 @56x56 layer
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 2 -c 1 \
    -f -o gpurun_out/prof_gemm_s8_win3x3_64x56 python bench_conv_layers.py --s8-out --only 2 > gpurun_out/ncu_win.log 2>&1
echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 2 -c 1 \
    -f -o gpurun_out/prof_gemm_s8_win3x3_128x28 python bench_conv_layers.py --s8-out --only 10 > gpurun_out/ncu_win2.log 2>&1
echo "rc=$?"
