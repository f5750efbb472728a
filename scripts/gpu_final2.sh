#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/z2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/z2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z2_smoke.log 2>&1; tail -1 gpurun_out/z2_smoke.log
for c in 3 5; do
  timeout 1500 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r02_bench_config$c.json 2> gpurun_out/z2_bench_c$c.err; echo "config $c rc=$?"
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_bench_config$c.json').read()); print({k:d[k] for k in ('value','ms_per_step','verified','gpu_launches')}, 'e2e', round(d['e2e']['value'],1), 'kernel_ms', d['roofline']['kernel_ms'], 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'cpu', round(d['cpu_baseline']['value'],3))" || tail -5 gpurun_out/z2_bench_c$c.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/z2_err.txt; python -c "import json; d=json.loads(open('gpurun_out/r02_bench_n1.json').read()); print('default', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['verified'])"
