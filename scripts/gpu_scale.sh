#!/bin/bash
# weak-scaling lines on one 8-GPU box: config 2 at N = 1, 2, 4, 8 (and config 5 at N = 8: 64 GiB in total)
mkdir -p gpurun_out
line() { python -c "import sys,json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'verified', d.get('verified'))" || tail -5 $3; }
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_c2_n1.json 2> gpurun_out/s_err1.txt; line gpurun_out/r02_scale_c2_n1.json "c2 n1" gpurun_out/s_err1.txt
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_c2_n$n.json 2> gpurun_out/s_err$n.txt; line gpurun_out/r02_scale_c2_n$n.json "c2 n$n" gpurun_out/s_err$n.txt
done
NCCL_DEBUG=INFO timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s_c5_n8.out 2> gpurun_out/s_err58.txt
grep -v "NCCL INFO" gpurun_out/s_c5_n8.out | tail -1 > gpurun_out/r02_scale_c5_n8.json; line gpurun_out/r02_scale_c5_n8.json "c5 n8" gpurun_out/s_err58.txt
grep -h "NCCL INFO" gpurun_out/s_c5_n8.out gpurun_out/s_err58.txt | grep -E "via|NVLS|Channel 00|Connected|Using network|P2P" | head -12 > gpurun_out/r02_nccl_transport.txt; cat gpurun_out/r02_nccl_transport.txt | cut -c1-200
rm -f gpurun_out/s_c5_n8.out
