#!/bin/bash
# gpurun wrapper (development): sim tests + headline bench without the CPU baseline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_intsim.py tests/test_gpu_int8_pipeline.py tests/test_gpu_e2e.py -m gpu -q -x > gpurun_out/pytest_sim.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sim.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/pytest_sim.log | tail -8
timeout 900 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
cat gpurun_out/bench_quick.json
timeout 900 python bench_sim.py --mode model > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "sim rc=$?"
cat gpurun_out/bench_sim.json
