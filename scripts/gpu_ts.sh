#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc_conv.py -m gpu -x -q -k "fused" 2>&1 | tail -5 | tee gpurun_out/pytest_pair.log
if grep -q "passed" gpurun_out/pytest_pair.log && ! grep -q "failed\|error" gpurun_out/pytest_pair.log; then
timeout 300 python scripts/prof_pair_ts.py 2>&1 | tee gpurun_out/prof_pair_ts.txt
rm -f gpurun_out/bench_pairs.txt
for P in off smem2; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-callers --no-sweep --no-configs --pairs $P 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$P', d['value'], d['ms_per_step'], d['stages_ms'], d['e2e']['value'], {k: round(v['ms'],3) for k,v in d['roofline_stages'].items()})" | tee -a gpurun_out/bench_pairs.txt
done
fi
