#!/bin/bash
# gpurun wrapper (round 2, experiment): periodic per-channel max-abs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "channel" > gpurun_out/pytest_chan.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_chan.log | cut -c1-200
timeout 600 python bench_stats.py --skip-stats --skip-fakequant > gpurun_out/bench_stats_chan.jsonl 2> gpurun_out/bench_stats_chan.err; echo "stats rc=$?"
cut -c1-220 gpurun_out/bench_stats_chan.jsonl
