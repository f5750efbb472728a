#!/bin/bash
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(open('$1').read()); print('$2', round(d['value'],1), 'GB/s', round(d['ms_per_step'],4), 'ms kernel', round(d['roofline']['kernel_ms'],4), d['scan_stats'].get('engine'), d.get('verified') is not None)" || tail -3 gpurun_out/q_err.txt; }
timeout -k 10 300 python scripts/sieve_check.py > gpurun_out/q_check.log 2>&1; echo "check rc=$?"; grep -c "^ok" gpurun_out/q_check.log; grep -E "^BAD|FAILURES|Error" gpurun_out/q_check.log | head -5
timeout -k 10 300 python bench.py --config 5 --scale 0.125 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_c5.json 2> gpurun_out/q_err.txt; show gpurun_out/q_c5.json "c5/8"
timeout -k 10 300 python bench.py --config 4 --scale 0.125 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_c4.json 2> gpurun_out/q_err.txt; show gpurun_out/q_c4.json "c4/8"
timeout -k 10 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_c3.json 2> gpurun_out/q_err.txt; show gpurun_out/q_c3.json "c3"
timeout -k 10 300 python bench.py --kernel 5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q_c2s.json 2> gpurun_out/q_err.txt; show gpurun_out/q_c2s.json "c2 sieve"
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sieve or not (staged or plain or global)" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q_pytest.log
