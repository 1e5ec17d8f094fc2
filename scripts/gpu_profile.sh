#!/bin/bash
# Profiling visit: launch list with DRAM bytes for one bench step + full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-callers > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_conv_kernel<\(int\)128, \(int\)0' -s 25 -c 1 -o gpurun_out/prof_tc128_bench python bench.py --steps 2 --warmup 3 --no-cpu --no-callers > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-160
ls -la gpurun_out/*.ncu-rep
