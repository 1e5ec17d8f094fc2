#!/bin/bash
# Final round visit: full GPU suite, smoke, headline bench, batch sweep, launch list of the final build.
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench.json
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['stages_ms'], d['e2e']['value'], d['roofline']['achieved'], d['cpu_baseline']['value'])"
for b in 1 8 128; do
  timeout 300 python bench.py --batch $b --steps 10 --no-cpu --no-callers 2>&1 | tail -1 > gpurun_out/bench_b$b.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_b$b.json')); print($b, d['value'], d['ms_per_step'], d['stages_ms'], d['e2e']['value'], d['rtf'])"
done
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-callers > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-80
