#!/bin/bash
# full GPU suite + smoke + single-GPU bench line (with cpu_baseline) + reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r3f_pytest.log 2>&1
tail -6 gpurun_out/r3f_pytest.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3f_smoke.log 2>&1; tail -2 gpurun_out/r3f_smoke.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r3f_bench_1gpu_c3.json 2> gpurun_out/r3f_bench.err; tail -2 gpurun_out/r3f_bench.err | cut -c1-200
python tools/stage_ms.py gpurun_out/r3f_bench_1gpu_c3.json
