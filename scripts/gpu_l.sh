#!/bin/bash
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(open('$1').read()); print('$2', round(d['value'],1), 'GB/s', round(d['ms_per_step'],4), 'ms kernel', round(d['roofline']['kernel_ms'],4), 'matches/s', '%.3g' % d['matches_per_s'], d['scan_stats'].get('engine'), d.get('verified') is not None)" || tail -3 gpurun_out/l_err.txt; }
for t in 8192 16384 32768 65536; do
  timeout 300 python bench.py --kernel 5 --segment-bytes $t --steps 20 --warmup 5 --no-cpu-baseline --no-verify > gpurun_out/l_c2_t$t.json 2> gpurun_out/l_err.txt; show gpurun_out/l_c2_t$t.json "c2 sieve T=$t"
done
for t in 8192 16384 32768 65536; do
  timeout 600 python bench.py --config 5 --scale 0.125 --kernel 5 --segment-bytes $t --steps 5 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/l_c5_t$t.json 2> gpurun_out/l_err.txt; show gpurun_out/l_c5_t$t.json "c5/8 sieve T=$t"
done
timeout 600 python bench.py --dense --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/l_c2_dense.json 2> gpurun_out/l_err.txt; show gpurun_out/l_c2_dense.json "c2 dense auto"
timeout 600 python bench.py --dense --kernel 2 --steps 10 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/l_c2_dense_k2.json 2> gpurun_out/l_err.txt; show gpurun_out/l_c2_dense_k2.json "c2 dense staged"
timeout 900 python bench.py --config 3 --dense --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/l_c3_dense.json 2> gpurun_out/l_err.txt; show gpurun_out/l_c3_dense.json "c3 dense auto"
timeout 300 python scripts/ragged_text.py > gpurun_out/l_ragged.log 2>&1; cat gpurun_out/l_ragged.log
ACB200_ENGINE=sieve timeout 300 python scripts/ragged_text.py > gpurun_out/l_ragged_sieve.log 2>&1; cat gpurun_out/l_ragged_sieve.log
