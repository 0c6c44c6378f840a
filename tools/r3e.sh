#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_two_level.py -q > gpurun_out/r3e_pytest.log 2>&1
tail -60 gpurun_out/r3e_pytest.log | cut -c1-260
