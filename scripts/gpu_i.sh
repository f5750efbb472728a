#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/i_pytest.log
for c in 2 3 4 5; do
  timeout 1500 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/i_bench_c${c}.json 2> gpurun_out/i_bench_c${c}.err; echo "config $c rc=$?"
  python -c "import sys,json; d=json.loads(open('gpurun_out/i_bench_c${c}.json').read()); print({k:d[k] for k in ('value','ms_per_step','e2e','verified','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['cpu_baseline']['value'])" || tail -5 gpurun_out/i_bench_c${c}.err
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.log 2>&1; tail -2 gpurun_out/i_smoke.log
