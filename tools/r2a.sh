#!/bin/bash
# round 2, first GPU call: the whole -m gpu suite, then the A-Trous forms on the C3 frame
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
for v in "tma 4" "tma 2" "cpasync 4" "legacy 4"; do
  set -- $v
  python bench.py --steps 16 --warmup 4 --no-cpu-baseline --denoiser $1 --tile-rows $2 > gpurun_out/r2a_bench_$1_$2.json 2> gpurun_out/r2a_bench_$1_$2.err
done
python tools/stage_ms.py gpurun_out/r2a_bench_*.json
cat gpurun_out/r2a_pytest.log
