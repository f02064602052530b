#!/bin/bash
# gpurun wrapper: ncu launch list of the bench command + full captures of this repo's kernels.
mkdir -p gpurun_out
export PQ_BENCH_NO_AUTOTUNE=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
for K in hist_multi absmax_multi kl_candidate fakequant; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 \
      -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "$K rc=$?"
done
# fake-quant is not on the calibration path: profile it from the GPU test that runs it at C2 scale
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fakequant_kernel -s 8 -c 1 \
    -f -o gpurun_out/prof_fakequant python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fakequant_full_size" > gpurun_out/ncu_fakequant.log 2>&1
echo "fakequant rc=$?"
ls -la gpurun_out/*.ncu-rep
