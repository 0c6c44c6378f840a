#!/bin/bash
mkdir -p gpurun_out
for tb in 0 740 592 444 296; do
  timeout 300 python bench.py --steps 32 --warmup 6 --no-cpu-baseline --frames-in-flight 2 --trace-blocks $tb > gpurun_out/r2f_tb$tb.json 2> gpurun_out/r2f_tb$tb.err
done
timeout 300 python bench.py --steps 32 --warmup 6 --no-cpu-baseline --frames-in-flight 1 --trace-blocks 444 > gpurun_out/r2f_fif1_tb444.json 2> gpurun_out/r2f_fif1.err
python tools/stage_ms.py gpurun_out/r2f_*.json
