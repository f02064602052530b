#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_intsim.py tests/test_gpu_c4_at_size.py tests/test_gpu_int8_pipeline.py -m gpu -q -x > gpurun_out/pytest_tps.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tps.log
tail -5 gpurun_out/pytest_tps.log
timeout 600 python bench_conv_layers.py --s8-out 2>&1 | tee gpurun_out/conv_layers_s8.txt | tail -28
for r in 3 7 13 19; do timeout 200 python bench_conv_layers.py --s8-out --fused-add --only $r 2>&1 | grep "^(" ; done | tee gpurun_out/fused_layers.txt
timeout 600 python bench_sim.py --mode model --iters 5 > gpurun_out/bench_sim.json 2> gpurun_out/bench_sim.err; echo "sim rc=$?"
python - <<PY
import json
for l in open("gpurun_out/bench_sim.json"):
    d=json.loads(l); print(d["config"]["variant"][:40], d["ms_per_forward"], {k:(v.get("ms_per_fwd"), v.get("launches_per_fwd")) for k,v in d["kernels"].items()})
PY
