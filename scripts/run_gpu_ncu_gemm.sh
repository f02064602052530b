#!/bin/bash
# gpurun wrapper: ncu --set full captures of the int8 GEMM/conv kernel on chosen rows of bench_conv_layers.py
mkdir -p gpurun_out
for ROW in ${ROWS:-0 1 2 3}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 3 -c 1 \
      -f -o gpurun_out/prof_gemm${TAG}_row$ROW python bench_conv_layers.py ${FLAGS---s8-out} --only $ROW > gpurun_out/ncu_gemm${TAG}_row$ROW.log 2>&1
  echo "row $ROW rc=$?"
done
ls -la gpurun_out/*.ncu-rep
