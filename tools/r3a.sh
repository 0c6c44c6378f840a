#!/bin/bash
# round-2 session 2: stage-pipeline parity on ONE GPU (all ranks share cuda:0 through CUDA IPC)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -q > gpurun_out/r3a_pytest.log 2>&1
tail -40 gpurun_out/r3a_pytest.log | cut -c1-250
