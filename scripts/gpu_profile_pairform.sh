#!/bin/bash
# ncu capture of the CTA-pair form of the N=128 conv (one launch)
mkdir -p gpurun_out
export VTTS_TC_VARIANT=3
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_conv_kernel<\(int\)128, \(int\)0, \(int\)2' -s 3 -c 1 -f -o gpurun_out/r2_tc_conv128_pairform python scripts/profile_driver.py step > gpurun_out/ncu_pairform.log 2>&1
tail -2 gpurun_out/ncu_pairform.log | cut -c1-150
ls -la gpurun_out/r2_tc_conv128_pairform.ncu-rep
