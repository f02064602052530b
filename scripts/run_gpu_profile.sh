#!/bin/bash
# gpurun wrapper: ncu launch list of the bench command + full captures of this repo's kernels.
mkdir -p gpurun_out
export PQ_BENCH_NO_AUTOTUNE=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
for K in hist_multi absmax_multi kl_candidate; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 \
      -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "$K rc=$?"
done
# fake-quant is not on the calibration path: profile it from the GPU test that runs it at C2 scale
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fakequant_kernel -s 0 -c 1 \
    -f -o gpurun_out/prof_fakequant python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fakequant_full_size" > gpurun_out/ncu_fakequant.log 2>&1
echo "fakequant rc=$?"
# int8 tensor-core kernel: one compute-bound 3x3 layer (row 16) and one store-bound 1x1 layer (row 3), int8 output
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_gemm_s8_conv3x3_256x14 python bench_conv_layers.py --s8-out --only 16 > gpurun_out/ncu_gemm1.log 2>&1
echo "gemm 3x3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_gemm_s8_conv1x1_64to256x56 python bench_conv_layers.py --s8-out --only 3 > gpurun_out/ncu_gemm2.log 2>&1
echo "gemm 1x1 rc=$?"
# the stem row kernel and the int8-pipeline add kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_rows_s8_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_conv_rows_stem python bench_conv_layers.py --s8-out --only 0 > gpurun_out/ncu_rows.log 2>&1
echo "rows rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:add_requant_kernel -s 20 -c 1 \
    -f -o gpurun_out/prof_add_requant python bench_sim.py --mode model --iters 1 > gpurun_out/ncu_add.log 2>&1
echo "add rc=$?"
ls -la gpurun_out/*.ncu-rep
unset PQ_BENCH_NO_AUTOTUNE
timeout 300 python scripts/exp_forward_variants.py > gpurun_out/exp_fwd.txt 2>&1; cat gpurun_out/exp_fwd.txt | tail -6
