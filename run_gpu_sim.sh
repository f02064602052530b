#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench_sim.py --mode both > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "rc=$?"
tail -5 gpurun_out/bench_sim.err | cut -c1-400; cat gpurun_out/bench_sim.json
