#!/bin/bash
# gpurun wrapper: fused-add tests + per-layer table + ResNet-50 forward, histogram families
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_int8_pipeline.py tests/test_gpu_parity.py tests/test_gpu_e2e.py -m gpu -q -x > gpurun_out/pytest_perf.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_perf.log
tail -5 gpurun_out/pytest_perf.log
for r in 3 7 13 19; do timeout 200 python bench_conv_layers.py --s8-out --fused-add --only $r 2>&1 | grep "^(" ; done | tee gpurun_out/fused_layers.txt
timeout 600 python bench_sim.py --mode model --iters 5 > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "sim rc=$?"
python - <<PY
import json
for l in open("gpurun_out/bench_sim.json"):
    d=json.loads(l); print(d["config"]["variant"][:40], d["ms_per_forward"], {k:(v.get("ms_per_fwd"), v.get("launches_per_fwd")) for k,v in d["kernels"].items()})
PY
timeout 900 python bench_stats.py --min-log2 28 --max-log2 32 --skip-channel --skip-fakequant > gpurun_out/stats.jsonl 2> gpurun_out/stats.err; echo "stats rc=$?"
python - <<PY
import json
for l in open("gpurun_out/stats.jsonl"):
    d=json.loads(l)
    if d.get("bench") == "stats": print(d["family"], d["log2_elements"], "hist_ms", d["hist_ms"], "frac", d["hist_frac"], "absmax_frac", d["absmax_frac"])
PY
