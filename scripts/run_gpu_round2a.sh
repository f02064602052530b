#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_intsim.py tests/test_gpu_parity.py tests/test_gpu_int8_pipeline.py -m gpu -q -x --durations=5 > gpurun_out/pytest_feat.log 2>&1; echo "feat rc=$?" >> gpurun_out/pytest_feat.log
tail -15 gpurun_out/pytest_feat.log
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench.time; echo "bench rc=$?"; grep "\[bench\]" gpurun_out/bench.err; tail -3 gpurun_out/bench.err | cut -c1-400; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.time
