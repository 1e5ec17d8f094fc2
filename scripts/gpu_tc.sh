#!/bin/bash
mkdir -p gpurun_out
echo "== tc probe"
timeout 180 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1; echo "probe rc=$?"
tail -70 gpurun_out/tc_probe.log
if grep -q PROBE_DONE gpurun_out/tc_probe.log; then
echo "== pytest tc"
timeout 600 python -m pytest tests/test_gpu_tc_conv.py -m gpu -x -q -s 2>&1 | tail -30 | tee gpurun_out/pytest_tc.log
fi
