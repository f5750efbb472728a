#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/host_path_timing.py > gpurun_out/h_host.log 2>&1; cat gpurun_out/h_host.log | tail -12
timeout 900 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu -k "pipeline or threads" > gpurun_out/h_pytest.log 2>&1; tail -3 gpurun_out/h_pytest.log
