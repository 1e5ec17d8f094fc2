#!/bin/bash
# same-box A/B of two builds: viettts_b200/lib_prev.bin (previous) vs viettts_b200/libviettts_b200.so (new)
mkdir -p gpurun_out
: > gpurun_out/ab_lib.txt
cp viettts_b200/libviettts_b200.so /tmp/new.so
run() {
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-callers --no-sweep --no-configs 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,2), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stages_ms'].items()}, {k.replace('hifigan_',''): round(v['ms'],3) for k,v in d['roofline_stages'].items() if 'hifigan' in k})" | tee -a gpurun_out/ab_lib.txt
}
for i in 1 2; do
  cp viettts_b200/lib_prev.bin viettts_b200/libviettts_b200.so; run prev
  cp /tmp/new.so viettts_b200/libviettts_b200.so; run new
done
[ -n "$TESTS" ] && timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
