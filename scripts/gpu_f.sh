#!/bin/bash
mkdir -p gpurun_out
for b in 1 2 4; do
  ACB200_EPILOGUE_BPS=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-verify > gpurun_out/f_bench_c2_bps$b.json 2> gpurun_out/f_err.txt
  python -c "import sys,json; d=json.loads(open('gpurun_out/f_bench_c2_bps$b.json').read()); print('bps$b', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])" || tail -3 gpurun_out/f_err.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err; echo "c2 rc=$?"; python -c "import sys,json; d=json.loads(open('gpurun_out/f_bench_c2.json').read()); print({k:d[k] for k in ('value','ms_per_step','roofline','e2e','verified','cpu_baseline','gpu_launches','scan_stats')})" || tail -3 gpurun_out/f_bench_c2.err
for c in 3 4 5; do
  timeout 1500 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/f_bench_c${c}.json 2> gpurun_out/f_bench_c${c}.err; echo "config $c rc=$?"
  python -c "import sys,json; d=json.loads(open('gpurun_out/f_bench_c${c}.json').read()); print({k:d[k] for k in ('value','ms_per_step','roofline','e2e','verified','cpu_baseline','gpu_launches')})" || tail -5 gpurun_out/f_bench_c${c}.err
done
timeout 1500 python -m pytest tests/test_gpu_full_size.py tests/golden/ref_tests -x -q -m gpu > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/f_pytest.log
