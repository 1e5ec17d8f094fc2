#!/bin/bash
mkdir -p gpurun_out
echo "== tc probe"
timeout 240 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1; echo "probe rc=$?"
grep -E "MISMATCH|WORST|PROBE_DONE|Error|error" gpurun_out/tc_probe.log | head -20
if grep -q PROBE_DONE gpurun_out/tc_probe.log; then
echo "== pytest tc"
timeout 600 python -m pytest tests/test_gpu_tc_conv.py -m gpu -x -q -s 2>&1 | tail -8 | tee gpurun_out/pytest_tc.log
echo "== stall counters"
timeout 300 python scripts/prof_conv.py 2>&1 | tail -14 | tee gpurun_out/prof_conv.log
echo "== bench"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['achieved'])"
fi
