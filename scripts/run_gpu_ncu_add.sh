#!/bin/bash
# gpurun wrapper: ncu --set full capture of the fused conv+add kernel on the 64->256 @56^2 layer (row 3) and the 512->2048 @7^2 layer (row 19)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_int8_pipeline.py -m gpu -q -x -k "fused or writes_only" 2>&1 | tail -3
for r in 3 19; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 2 -c 1 -f -o gpurun_out/prof_fused_add_row$r \
    python bench_conv_layers.py --s8-out --fused-add --only $r > gpurun_out/ncu_fused_row$r.log 2>&1; echo "ncu row $r rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -4
