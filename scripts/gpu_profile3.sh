#!/bin/bash
# Round-2 ncu captures of the kernels that changed after scripts/gpu_profile2.sh (CTA-pair conv, multi-row-group decoder
# scan, cp.async teacher-forced scan) plus the two ConvTranspose kernels; --set full, one launch each.
mkdir -p gpurun_out
cap() {  # name mode regex skip
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$3" -s $4 -c 1 -f -o gpurun_out/r2_$1 python scripts/profile_driver.py $2 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-120
}
cap tc_conv128 step 'tc_conv_kernel<\(int\)128, \(int\)0, \(int\)2' 3
cap tc_conv256 step 'tc_conv_kernel<\(int\)256, \(int\)0' 5
cap ups64      step 'tc_conv_kernel<\(int\)64, \(int\)0, \(int\)2, \(int\)2' 0
cap ups32      step 'tc_conv_kernel<\(int\)32, \(int\)0, \(int\)4, \(int\)2' 0
cap decoder_scan step 'decoder_scan_kernel' 0
cap tf_scan    gta  'decoder_tf_scan_kernel' 0
ls -la gpurun_out/r2_*.ncu-rep | awk '{print $5, $9}'
