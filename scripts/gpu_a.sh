#!/bin/bash
# round-2 GPU session A: first contact of the sieve kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 600 python scripts/sieve_check.py > gpurun_out/a_check.log 2>&1; echo "check rc=$?" >> gpurun_out/a_check.log
tail -5 gpurun_out/a_check.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sieve_check.py > gpurun_out/a_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/a_memcheck.log
grep -E "ERROR SUMMARY|Invalid|rc=" gpurun_out/a_memcheck.log | head -5
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sieve or not (staged or plain or global)" > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench_sieve.json 2> gpurun_out/a_bench_sieve.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel 2 > gpurun_out/a_bench_staged.json 2> gpurun_out/a_bench_staged.err
cat gpurun_out/a_bench_sieve.json gpurun_out/a_bench_staged.json | cut -c1-600
timeout 600 python scripts/other_configs.py > gpurun_out/a_other.log 2>&1
cat gpurun_out/a_other.log
timeout 300 python scripts/ragged_text.py > gpurun_out/a_ragged.log 2>&1
cat gpurun_out/a_ragged.log
