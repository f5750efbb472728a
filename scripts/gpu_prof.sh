#!/bin/bash
# round-2 profiles: ncu --set full of the dominant kernel per config + the launch list of the default bench
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_staged -s 3 -c 1 -o gpurun_out/r02_prof_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sieve_scan -s 3 -c 1 -o gpurun_out/r02_prof_c2_sieve python bench.py --kernel 5 --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c2s.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sieve_scan -s 2 -c 1 -o gpurun_out/r02_prof_c3 python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c3.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sieve_scan -s 2 -c 1 -o gpurun_out/r02_prof_c5 python bench.py --config 5 --steps 1 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c5.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sieve_scan -s 2 -c 1 -o gpurun_out/r02_prof_c4 python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sieve_epilogue -s 2 -c 1 -o gpurun_out/r02_prof_c3_epilogue python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c3e.log 2>&1
ls -la gpurun_out/r02_*
