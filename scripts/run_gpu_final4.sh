#!/bin/bash
# gpurun wrapper (round 2, final evidence, slim edition: gpurun merges at most 64 MiB back, so only four ncu captures;
# the GPU test suite of the same tree is in profiles/r02_final_run_stdout.log: 327 passed).
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; cat gpurun_out/bench_ref.json | cut -c1-300
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench.time; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/bench.err; tail -3 gpurun_out/bench.time
timeout 900 python bench_sim.py --mode both > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "sim rc=$?"
timeout 600 python bench_conv_layers.py > gpurun_out/conv_layers.txt 2>&1
timeout 600 python bench_conv_layers.py --s8-out > gpurun_out/conv_layers_s8.txt 2>&1
tail -3 gpurun_out/conv_layers_s8.txt
for r in 2 10; do timeout 200 python bench_conv_layers.py --s8-out --no-windows --only $r 2>&1 | grep "^(" ; done | tee gpurun_out/conv_layers_no_windows.txt
for r in 3 7 13 19; do timeout 200 python bench_conv_layers.py --s8-out --fused-add --only $r 2>&1 | grep "^(" ; done | tee gpurun_out/fused_layers.txt
timeout 900 python bench_stats.py > gpurun_out/bench_stats.jsonl 2> gpurun_out/bench_stats.err; echo "stats rc=$?"
export PQ_BENCH_NO_AUTOTUNE=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
unset PQ_BENCH_NO_AUTOTUNE
cap() {   # name, extra args of bench_conv_layers.py
    local name=$1; shift
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_s8_kernel -s 2 -c 1 \
        -f -o gpurun_out/prof_$name python bench_conv_layers.py --s8-out "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "$name rc=$?"
}
cap gemm_s8_win3x3_64x56 --only 2
cap gemm_s8_win3x3_128x28 --only 10
cap fused_add_64to256x56 --fused-add --only 3
cap fused_add_256to1024x14 --fused-add --only 13
du -sh gpurun_out
