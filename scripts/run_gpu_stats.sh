#!/bin/bash
# gpurun wrapper: GPU tests, the C5 statistics microbenchmark (bench_stats.py), a short headline bench and ncu
# captures of the kernels that have no summary under profiles/ yet.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench_stats.py --iters 5 > gpurun_out/bench_stats.jsonl 2> gpurun_out/bench_stats.err; echo "bench_stats rc=$?"
tail -3 gpurun_out/bench_stats.err | cut -c1-300; cat gpurun_out/bench_stats.jsonl | cut -c1-400
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err | cut -c1-300; cat gpurun_out/bench.json
# ncu: per-channel max-abs, histogram on the single-bin worst case, fake-quant, add_requant, stem row kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:absmax_per_channel -s 0 -c 1 \
    -f -o gpurun_out/prof_absmax_per_channel python -m pytest tests/test_gpu_parity.py -q -m gpu -k "per_channel_vs_oracle and 32x256" > gpurun_out/ncu_chan.log 2>&1
echo "chan rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hist_multi -s 3 -c 1 \
    -f -o gpurun_out/prof_hist_const python bench_stats.py --families const --min-log2 28 --max-log2 28 --skip-fakequant > gpurun_out/ncu_hist_const.log 2>&1
echo "hist const rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fakequant_kernel -s 0 -c 1 \
    -f -o gpurun_out/prof_fakequant python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fakequant_full_size" > gpurun_out/ncu_fakequant.log 2>&1
echo "fakequant rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_rows_s8_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_conv_rows_stem python bench_conv_layers.py --s8-out --only 0 > gpurun_out/ncu_rows.log 2>&1
echo "rows rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:add_requant_kernel -s 20 -c 1 \
    -f -o gpurun_out/prof_add_requant python bench_sim.py --mode model --iters 1 > gpurun_out/ncu_add.log 2>&1
echo "add rc=$?"
ls -la gpurun_out/*.ncu-rep
