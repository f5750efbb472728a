#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 8 --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c5n8.out 2> gpurun_out/c5n8_err.txt; tail -1 gpurun_out/c5n8.out > gpurun_out/r02_scale_c5_n8.json; python -c "import sys,json; d=json.loads(open('gpurun_out/r02_scale_c5_n8.json').read()); print('c5 n8', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['verified'], d['e2e']['value'])" || tail -5 gpurun_out/c5n8_err.txt
