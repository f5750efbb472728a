#!/bin/bash
# Round-2 profiles (one GPU): launch list of the default bench, ncu --set full of the dominant kernel per config,
# summarised ON THE BOX (gpurun_out/ may hold 64 MiB: the 16 MB reports are deleted after summarising, two are kept).
mkdir -p gpurun_out
sum() { python profiles/summarize.py gpurun_out/$1.ncu-rep $2 gpurun_out/$3 > /dev/null 2>> gpurun_out/p_sum.err || echo "summarize $1 failed"; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_staged -s 3 -c 1 -o gpurun_out/r02_prof_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c2.log 2>&1
sum r02_prof_c2 scan_staged_kernel r02_config2_scan_kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:epilogue_kernel -s 3 -c 1 -o gpurun_out/r02_prof_c2_epi python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c2e.log 2>&1
sum r02_prof_c2_epi epilogue_kernel r02_config2_epilogue_kernel; rm -f gpurun_out/r02_prof_c2_epi.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sieve_scan -s 3 -c 1 -o gpurun_out/r02_prof_c2_sieve python bench.py --kernel 5 --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c2s.log 2>&1
sum r02_prof_c2_sieve sieve_scan_kernel r02_config2_sieve_scan_kernel; rm -f gpurun_out/r02_prof_c2_sieve.ncu-rep
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sieve_scan -s 2 -c 1 -o gpurun_out/r02_prof_c3 python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c3.log 2>&1
sum r02_prof_c3 sieve_scan_kernel r02_config3_scan_kernel; rm -f gpurun_out/r02_prof_c3.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sieve_epilogue -s 2 -c 1 -o gpurun_out/r02_prof_c3_epi python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c3e.log 2>&1
sum r02_prof_c3_epi sieve_epilogue_kernel r02_config3_epilogue_kernel; rm -f gpurun_out/r02_prof_c3_epi.ncu-rep
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sieve_scan -s 2 -c 1 -o gpurun_out/r02_prof_c5 python bench.py --config 5 --steps 1 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c5.log 2>&1
sum r02_prof_c5 sieve_scan_kernel r02_config5_scan_kernel; rm -f gpurun_out/r02_prof_c5.ncu-rep
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sieve_scan -s 2 -c 1 -o gpurun_out/r02_prof_c4 python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/p_c4.log 2>&1
sum r02_prof_c4 sieve_scan_kernel r02_config4_scan_kernel
rm -f gpurun_out/*.log.tmp; du -sh gpurun_out; ls gpurun_out | head -40
