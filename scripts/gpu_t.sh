#!/bin/bash
mkdir -p gpurun_out
ACB200_TRACE=1 timeout -k 10 900 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/t_c5.log 2> gpurun_out/t_err.txt; grep "trace" gpurun_out/t_c5.log | tail -12; tail -1 gpurun_out/t_c5.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"; tail -3 gpurun_out/t_err.txt
timeout -k 10 900 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/t_c3.json 2>> gpurun_out/t_err.txt; python -c "import json; d=json.loads(open('gpurun_out/t_c3.json').read()); print('c3', d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'])"
