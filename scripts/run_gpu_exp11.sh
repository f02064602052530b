#!/bin/bash
# gpurun wrapper (round 2, experiment): space-to-depth stem
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_intsim.py tests/test_gpu_int8_pipeline.py tests/test_gpu_lenet.py tests/test_gpu_parity.py -m gpu -x -q -k "smallc or pipeline_equals or lenet or golden or abi" > gpurun_out/pytest_dev.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_dev.log | cut -c1-200
timeout 300 python bench_conv_layers.py --s8-out --only 0 2>&1 | grep "^("
timeout 300 python bench_conv_layers.py --s8-out --only 0 --no-s2d 2>&1 | grep "^("
timeout 600 python bench_sim.py --mode model > gpurun_out/bench_sim_dev.json 2> gpurun_out/bench_sim_dev.err; echo "sim rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/bench_sim_dev.json"):
    d = json.loads(l)
    print(d["config"]["variant"][:60], d["ms_per_forward"], {k: round(v["ms_per_fwd"], 3) for k, v in d.get("kernels", {}).items()})
PY
