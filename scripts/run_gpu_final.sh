#!/bin/bash
# gpurun wrapper: everything the round-end driver runs (smoke, GPU tests, both bench arms) plus the secondary
# benches and the ncu evidence that goes under profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${STEPS:-16} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -2 gpurun_out/bench.err | cut -c1-300; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 900 python bench_sim.py --mode both > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "sim rc=$?"
cat gpurun_out/bench_sim.json | cut -c1-400
timeout 600 python bench_conv_layers.py > gpurun_out/conv_layers.txt 2>&1
timeout 600 python bench_conv_layers.py --s8-out > gpurun_out/conv_layers_s8.txt 2>&1
tail -3 gpurun_out/conv_layers_s8.txt
timeout 600 python bench_stats.py --skip-stats --skip-fakequant > gpurun_out/bench_stats_channel.jsonl 2> gpurun_out/bench_stats_channel.err
cat gpurun_out/bench_stats_channel.jsonl
export PQ_BENCH_NO_AUTOTUNE=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_multi -s 1 -c 1 \
    -f -o gpurun_out/prof_hist_multi $CMD > gpurun_out/ncu_hist.log 2>&1; echo "hist rc=$?"
unset PQ_BENCH_NO_AUTOTUNE
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_gemm_s8_conv3x3_256x14 python bench_conv_layers.py --s8-out --only 16 > gpurun_out/ncu_gemm1.log 2>&1
echo "gemm 3x3 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_gemm_s8_conv1x1_64to256x56 python bench_conv_layers.py --s8-out --only 3 > gpurun_out/ncu_gemm2.log 2>&1
echo "gemm 1x1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_gemm_s8_conv3x3_64x56 python bench_conv_layers.py --s8-out --only 2 > gpurun_out/ncu_gemm3.log 2>&1
echo "gemm 3x3 c64 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_rows_s8_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_conv_rows_stem python bench_conv_layers.py --s8-out --only 0 > gpurun_out/ncu_rows.log 2>&1
echo "rows rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:add_requant_kernel -s 20 -c 1 \
    -f -o gpurun_out/prof_add_requant python bench_sim.py --mode model --iters 1 > gpurun_out/ncu_add.log 2>&1
echo "add rc=$?"
ls gpurun_out/*.ncu-rep
