#!/bin/bash
# gpurun wrapper: fused conv+add development loop
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_int8_pipeline.py -m gpu -q -x -k "fused or pipeline_equals" > gpurun_out/pytest_add.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_add.log
tail -12 gpurun_out/pytest_add.log
for f in 0 1; do
  PQ_FUSE_ADD=$f timeout 600 python bench_sim.py --mode model --iters 5 > gpurun_out/bench_sim_fuse$f.json 2> gpurun_out/bench_sim_fuse$f.err; echo "sim fuse=$f rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/bench_sim_fuse$f.json"):
    d=json.loads(l); print(d["config"]["variant"][:40], d["ms_per_forward"], {k:(v.get("ms_per_fwd"), v.get("launches_per_fwd")) for k,v in d["kernels"].items()})
PY
done
