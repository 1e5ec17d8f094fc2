for V in 1 2; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-callers --no-sweep --no-configs --pairs smem2 --tc-variant $V 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $V', d['value'], d['ms_per_step'], d['stages_ms'], {k: round(v['ms'],3) for k,v in d['roofline_stages'].items() if 'hifigan' in k})" | tee -a gpurun_out/bench_variant.txt
done
timeout 300 python -m pytest tests/test_gpu_hifigan.py tests/test_gpu_tc_conv.py -m gpu -x -q 2>&1 | tail -3
