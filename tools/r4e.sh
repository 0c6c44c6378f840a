#!/bin/bash
# warm-cache launch list (ncu --cache-control none: durations as in the running pipeline, serialised)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/r4e_launches_warm.csv python bench.py --steps 2 --warmup 6 --no-cpu-baseline > gpurun_out/r4e.log 2>&1
tail -2 gpurun_out/r4e.log | cut -c1-300
