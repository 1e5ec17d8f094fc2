#!/bin/bash
# Round-2 ncu captures (--set full, one launch each): the tcgen05 generator kernels, the scans, the streaming kernels.
mkdir -p gpurun_out
cap() {  # name mode regex skip
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$3" -s $4 -c 1 -f -o gpurun_out/r2_$1 python scripts/profile_driver.py $2 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-120
}
cap pair2_64   step 'tc_pair2_kernel<\(int\)64' 1
cap pair2_32   step 'tc_pair2_kernel<\(int\)32' 1
cap tc_conv128 step 'tc_conv_kernel<\(int\)128, \(int\)0, \(int\)2' 3
cap tc_conv256 step 'tc_conv_kernel<\(int\)256, \(int\)0' 5
cap decoder_scan step 'decoder_scan_kernel' 0
cap enc_scan   step 'enc_scan_kernel' 0
cap conv_post  step 'conv_post_kernel' 0
cap upsample   step 'upsample_kernel' 0
cap tf_scan    gta  'decoder_tf_scan_kernel' 0
cap melspec    melspec 'melspec_kernel' 0
cap conv1d_fp32 fp32 'conv1d_nwc_kernel' 12
ls -la gpurun_out/r2_*.ncu-rep | awk '{print $5, $9}'
