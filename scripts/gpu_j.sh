#!/bin/bash
# two GPUs: NCCL parity test, weak-scaling bench lines
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu -k "nccl" > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/j_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/j_bench_c2_n2.json 2> gpurun_out/j_bench_c2_n2.err; echo "c2 n2 rc=$?"
python -c "import sys,json; d=json.loads(open('gpurun_out/j_bench_c2_n2.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','verified','gpu_launches','host_enqueue_ms_per_step')})" || tail -8 gpurun_out/j_bench_c2_n2.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-verify > gpurun_out/j_bench_c2_n1.json 2> gpurun_out/j_err.txt
python -c "import sys,json; d=json.loads(open('gpurun_out/j_bench_c2_n1.json').read()); print('n1', d['value'], d['ms_per_step'])"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_c5_n2.json 2> gpurun_out/j_bench_c5_n2.err; echo "c5 n2 rc=$?"
python -c "import sys,json; d=json.loads(open('gpurun_out/j_bench_c5_n2.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus','verified')})" || tail -8 gpurun_out/j_bench_c5_n2.err
