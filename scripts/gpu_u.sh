#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 scripts/nccl_gather_timing.py > gpurun_out/u_gather.log 2>&1; grep "world" gpurun_out/u_gather.log; tail -3 gpurun_out/u_gather.log | cut -c1-300
ACB200_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --config 5 --scale 0.5 --steps 3 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/u_c5n2.log 2> gpurun_out/u_err.txt; grep -c trace gpurun_out/u_c5n2.log; tail -1 gpurun_out/u_c5n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5/2 n2', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
