#!/bin/bash
# gpurun wrapper: both bench arms as the driver runs them + the new at-size tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_c4_at_size.py tests/test_gpu_intsim.py::test_accumulator_range_debug_check -m gpu -q --durations=5 > gpurun_out/pytest_c4.log 2>&1; echo "c4 rc=$?" >> gpurun_out/pytest_c4.log
tail -12 gpurun_out/pytest_c4.log
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; echo "ref rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.time
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench.time; echo "bench rc=$?"; tail -5 gpurun_out/bench.err | cut -c1-400; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.time
