#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 scripts/nccl_gather_timing.py > gpurun_out/v_gather.log 2>&1; grep "world" gpurun_out/v_gather.log
ACB200_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --config 5 --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/v_c5n8.log 2> gpurun_out/v_err.txt; grep "trace\] rank 0" gpurun_out/v_c5n8.log | tail -4; grep "trace\] run" gpurun_out/v_c5n8.log | tail -3; tail -1 gpurun_out/v_c5n8.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 n8', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
