#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sieve_check.py > gpurun_out/d_check.log 2>&1; echo "check rc=$?" >> gpurun_out/d_check.log
grep -c "^ok" gpurun_out/d_check.log; grep -E "^BAD|FAILURES|rc=|Error|error" gpurun_out/d_check.log | head
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench_sieve.json 2> gpurun_out/d_bench_sieve.err
cut -c1-420 gpurun_out/d_bench_sieve.json; tail -3 gpurun_out/d_bench_sieve.err
timeout 600 python scripts/other_configs.py > gpurun_out/d_other.log 2>&1; cat gpurun_out/d_other.log | cut -c1-200
timeout 300 python scripts/ragged_text.py > gpurun_out/d_ragged.log 2>&1; cat gpurun_out/d_ragged.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sieve_scan -s 3 -c 1 -o gpurun_out/d_prof_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/d_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sieve_scan -s 2 -c 1 -o gpurun_out/d_prof_c4 python scripts/other_configs.py c4 --short > gpurun_out/d_ncu_c4.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sieve or not (staged or plain or global)" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -4 gpurun_out/d_pytest.log
