#!/bin/bash
# gpurun wrapper (round 2, final evidence): what the round-end driver runs (smoke, the whole GPU test suite, both bench
# arms with the driver's flags) + the secondary benches; two small ncu captures (gpurun merges at most 64 MiB back).
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | tail -6
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; cat gpurun_out/bench_ref.json | cut -c1-200
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench.time; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/bench.err; tail -3 gpurun_out/bench.time
timeout 900 python bench_sim.py --mode both > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "sim rc=$?"
timeout 600 python bench_conv_layers.py > gpurun_out/conv_layers.txt 2>&1
timeout 600 python bench_conv_layers.py --s8-out > gpurun_out/conv_layers_s8.txt 2>&1
tail -3 gpurun_out/conv_layers_s8.txt
for r in 3 7 13 19; do timeout 200 python bench_conv_layers.py --s8-out --fused-add --only $r 2>&1 | grep "^(" ; done | tee gpurun_out/fused_layers.txt
timeout 900 python bench_stats.py > gpurun_out/bench_stats.jsonl 2> gpurun_out/bench_stats.err; echo "stats rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_rows_s8_kernel -s 2 -c 1 \
    -f -o gpurun_out/prof_conv_rows_stem python bench_conv_layers.py --s8-out --only 0 > gpurun_out/ncu_rows.log 2>&1; echo "rows rc=$?"
du -sh gpurun_out
