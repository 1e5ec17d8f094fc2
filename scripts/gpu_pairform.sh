#!/bin/bash
# CTA-pair form of the tensor-core conv (tc_variant 3): parity tests, then A/B bench against the default variant.
mkdir -p gpurun_out
export VTTS_TC_VARIANT=3
echo "== pytest (pair form)"
timeout 600 python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_hifigan.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_pairform.log
unset VTTS_TC_VARIANT
: > gpurun_out/bench_pairform.txt
for V in 1 3 1 3; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-callers --no-sweep --no-configs --tc-variant $V 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $V', d['value'], d['ms_per_step'], d['stages_ms'], {k: round(v['ms'],3) for k,v in d['roofline_stages'].items() if 'hifigan' in k})" | tee -a gpurun_out/bench_pairform.txt
done
