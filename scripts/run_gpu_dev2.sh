#!/bin/bash
# gpurun wrapper (round 2, short dev loop): fused-add parity tests + fused layer timings + bench_sim pipeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_int8_pipeline.py -m gpu -x -q > gpurun_out/pytest_dev.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_dev.log | cut -c1-200
for r in 3 7 13 19; do timeout 200 python bench_conv_layers.py --s8-out --fused-add --only $r 2>&1 | grep "^(" ; done | tee gpurun_out/fused_layers_dev.txt
timeout 600 python bench_sim.py --mode model > gpurun_out/bench_sim_dev.json 2> gpurun_out/bench_sim_dev.err; echo "sim rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/bench_sim_dev.json"):
    d = json.loads(l)
    print(d["config"]["variant"][:60], d["ms_per_forward"], {k: round(v["ms_per_fwd"], 3) for k, v in d.get("kernels", {}).items()})
PY
