#!/bin/bash
mkdir -p gpurun_out
for c in 2 3 4 5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 --scale 0.1 --no-cpu-baseline > gpurun_out/e_bench_c${c}_small.json 2> gpurun_out/e_bench_c${c}_small.err
  echo "config $c small rc=$?"; cut -c1-300 gpurun_out/e_bench_c${c}_small.json; tail -3 gpurun_out/e_bench_c${c}_small.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench_c2.json 2> gpurun_out/e_bench_c2.err; echo "c2 rc=$?"; cat gpurun_out/e_bench_c2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','roofline','e2e','verified','cpu_baseline','gpu_launches','scan_stats')})"; tail -3 gpurun_out/e_bench_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err; cut -c1-300 gpurun_out/e_bench_ref.json
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/e_pytest.log
