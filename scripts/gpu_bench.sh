#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS} 2>&1 | tail -2 | tee gpurun_out/bench.log
if [ -n "$NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu ${BENCH_ARGS} > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-300
fi
