#!/bin/bash
# one validation pass of the restructured sieve kernel: tests first, then the bench lines of the sieve configurations
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s_pytest.log
timeout 120 python scripts/sieve_check.py > gpurun_out/s_sieve_check.log 2>&1; echo "sieve_check rc=$?"; tail -2 gpurun_out/s_sieve_check.log
for c in 3 5 4; do
  timeout 150 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/s_bench_c$c.json 2> gpurun_out/s_bench_c$c.err; echo "config $c rc=$?"
  python -c "import sys,json; d=json.loads(open('gpurun_out/s_bench_c$c.json').read()); print({k:d[k] for k in ('value','ms_per_step','verified','gpu_launches')}, 'e2e', round(d['e2e']['value'],1), 'kernel_ms', d['roofline']['kernel_ms'], 'frac', round(d['roofline']['frac'],4))" || tail -5 gpurun_out/s_bench_c$c.err
done
ACB200_ENGINE=sieve timeout 100 python bench.py --config 2 --steps 20 --warmup 5 > gpurun_out/s_bench_c2_sieve.json 2> gpurun_out/s_bench_c2s.err; python -c "import json; d=json.loads(open('gpurun_out/s_bench_c2_sieve.json').read()); print('c2 sieve', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['verified'])"
