#!/bin/bash
# gpurun wrapper (round 2, experiment): programmatic dependent launch on / off -- parity tests, then bench_sim both ways
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_intsim.py tests/test_gpu_int8_pipeline.py tests/test_gpu_c4_at_size.py -m gpu -x -q > gpurun_out/pytest_dev.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_dev.log | cut -c1-200
for v in 0 1 0 1; do
  PQ_NO_PDL=$v timeout 600 python bench_sim.py --mode model > gpurun_out/bench_sim_pdl$v.json 2> gpurun_out/bench_sim_pdl$v.err; echo "no_pdl=$v rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/bench_sim_pdl$v.json"):
    d = json.loads(l)
    print(d["config"]["variant"][:60], d["ms_per_forward"], {k: round(x["ms_per_fwd"], 3) for k, x in d.get("kernels", {}).items()})
PY
done
