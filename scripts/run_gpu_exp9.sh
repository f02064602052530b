#!/bin/bash
# gpurun wrapper (round 2): ncu captures of the pipeline's glue kernels (stem quantiser, max-pool, average pool)
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
for k in maxpool_nhwc_s8 quantize_pad_nhwc8 avgpool_global; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 \
    -f -o gpurun_out/prof_$k python bench_sim.py --mode model --iters 2 > gpurun_out/ncu_$k.log 2>&1
echo "$k rc=$?"
done
du -sh gpurun_out
