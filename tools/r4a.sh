#!/bin/bash
# build-quality A/B: new tests first, then the bench with the SAH tree (default) and the Morton tree, then the full GPU suite under the new default
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_build_quality.py -x -q -m gpu ) > gpurun_out/r4a_pytest_bq.log 2>&1
tail -4 gpurun_out/r4a_pytest_bq.log | cut -c1-300
timeout 400 python bench.py --steps 32 --warmup 8 --no-cpu-baseline > gpurun_out/r4a_bench_sah.json 2> gpurun_out/r4a_bench_sah.err; tail -2 gpurun_out/r4a_bench_sah.err | cut -c1-200
python tools/stage_ms.py gpurun_out/r4a_bench_sah.json
EIDOLA_ACCEL_BUILD=lbvh timeout 400 python bench.py --steps 32 --warmup 8 --no-cpu-baseline > gpurun_out/r4a_bench_lbvh.json 2> gpurun_out/r4a_bench_lbvh.err; tail -2 gpurun_out/r4a_bench_lbvh.err | cut -c1-200
python tools/stage_ms.py gpurun_out/r4a_bench_lbvh.json
( time timeout 1200 python -m pytest tests/ -x -q -m gpu --deselect tests/test_gpu_build_quality.py ) > gpurun_out/r4a_pytest.log 2>&1
tail -4 gpurun_out/r4a_pytest.log | cut -c1-300
