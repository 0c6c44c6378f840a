#!/bin/bash
# usage: r3c.sh <N> "<label>|<bench args>" ...   (pipeline parity tests on cuda:0 first)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -x > gpurun_out/r3c_pytest.log 2>&1
tail -15 gpurun_out/r3c_pytest.log | cut -c1-250
bash tools/r3d.sh "$@"
