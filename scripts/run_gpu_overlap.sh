#!/bin/bash
mkdir -p gpurun_out
for ov in 0 1 0 1; do
  PQ_OVERLAP_STATS=$ov timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_ov$ov.json 2> gpurun_out/bench_ov$ov.err; echo "overlap=$ov rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_ov$ov.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'hist frac', d['roofline']['frac'], 'absmax', d['roofline']['absmax_GBps'], d['phases_s'])"
done
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | tail -12
