#!/bin/bash
# gpurun wrapper (round 2, experiment): histogram redo shortcut (statistics parity tests + C5 outlier / relu families)
# and the per-batch trace of the full-size config-3 job (scripts/exp_c3_full.py).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_stats.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_stats.log
timeout 600 python bench_stats.py --families outlier,relu,dense --min-log2 28 --max-log2 30 --skip-fakequant --skip-channel \
    > gpurun_out/bench_stats_exp.jsonl 2> gpurun_out/bench_stats_exp.err; echo "stats rc=$?"
cut -c1-260 gpurun_out/bench_stats_exp.jsonl
timeout 900 python scripts/exp_c3_full.py --both > gpurun_out/exp_c3_full.log 2>&1; echo "c3 rc=$?"
grep -v "^\s*$" gpurun_out/exp_c3_full.log | tail -45
