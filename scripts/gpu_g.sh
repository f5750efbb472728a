#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/g_pytest.log
for c in 2 3 5 4; do
  timeout 1500 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_c${c}.json 2> gpurun_out/g_bench_c${c}.err; echo "config $c rc=$?"
  python -c "import sys,json; d=json.loads(open('gpurun_out/g_bench_c${c}.json').read()); print({k:d[k] for k in ('value','ms_per_step','e2e','verified','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'])" || tail -5 gpurun_out/g_bench_c${c}.err
done
ACB200_ENGINE=sieve timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g_bench_c2_sieve.json 2> gpurun_out/g_err.txt
python -c "import sys,json; d=json.loads(open('gpurun_out/g_bench_c2_sieve.json').read()); print('c2 sieve', {k:d[k] for k in ('value','ms_per_step','e2e','verified')}, d['roofline']['kernel_ms'])" || tail -3 gpurun_out/g_err.txt
