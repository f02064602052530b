#!/bin/bash
# gpurun wrapper (round 2, experiment): the full-size config-3 job INSIDE bench.py (after the headline / e2e jobs), with
# the per-batch pass-1 trace and allocator statistics, twice in a row.
mkdir -p gpurun_out
PQ_BENCH_ONLY_C3=1 PQ_BENCH_C3_TRACE=1 PQ_BENCH_C3_REPEAT=2 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline \
    > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/bench_c3.err | cut -c1-260
