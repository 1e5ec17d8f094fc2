#!/bin/bash
for E in ${EXPS:-0}; do
echo "=== VTTS_PAIR_EXP=$E"
VTTS_PAIR_EXP=$E timeout 100 python scripts/prof_pairform.py 2>&1 | grep -v "rank 1"
done 2>&1 | tee gpurun_out/prof_pairexp.txt
[ -n "$SKIPTEST" ] || timeout 600 python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_hifigan.py -m gpu -x -q 2>&1 | tail -3
[ -n "$SKIPTEST" ] || VTTS_TC_VARIANT=3 timeout 600 python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_hifigan.py -m gpu -x -q 2>&1 | tail -3
: > gpurun_out/bench_pairform.txt
for cfg in "0 1" "0 3" "0 1" "0 3"; do
set -- $cfg
VTTS_PAIR_EXP=$1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-callers --no-sweep --no-configs --tc-variant $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('exp $1 variant $2', d['value'], d['ms_per_step'], d['stages_ms'], {k: round(v['ms'],3) for k,v in d['roofline_stages'].items() if 'hifigan' in k})" | tee -a gpurun_out/bench_pairform.txt
done
