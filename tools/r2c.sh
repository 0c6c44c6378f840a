#!/bin/bash
# GPU tests + 1-GPU bench (static and orbit)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest.log
python bench.py --steps 16 --warmup 4 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
python bench.py --steps 16 --warmup 4 --no-cpu-baseline --orbit 0.5 > gpurun_out/r2c_bench_n1_orbit.json 2> gpurun_out/r2c_bench_n1_orbit.err
python tools/stage_ms.py gpurun_out/r2c_bench_n1.json gpurun_out/r2c_bench_n1_orbit.json
tail -3 gpurun_out/r2c_bench_n1.err
cat gpurun_out/r2c_pytest.log
