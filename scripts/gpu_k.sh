#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/k_pytest.log
timeout 600 python scripts/host_path_timing.py > gpurun_out/k_host.log 2>&1; tail -9 gpurun_out/k_host.log
timeout 900 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_c4.json 2> gpurun_out/k_bench_c4.err; python -c "import sys,json; d=json.loads(open('gpurun_out/k_bench_c4.json').read()); print('c4', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])" || tail -5 gpurun_out/k_bench_c4.err
bash scripts/gpu_prof.sh
