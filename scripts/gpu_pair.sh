#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_conv.py -m gpu -x -q -k "fused" 2>&1 | tail -15 | tee gpurun_out/pytest_pair.log
if grep -q "passed" gpurun_out/pytest_pair.log && ! grep -q "failed\|error" gpurun_out/pytest_pair.log; then
timeout 400 python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_hifigan.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['e2e']['value'])"
fi
