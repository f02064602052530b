#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench_sim.py --mode ${MODE:-model} > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "rc=$?"
tail -3 gpurun_out/bench_sim.err | cut -c1-400; cat gpurun_out/bench_sim.json
