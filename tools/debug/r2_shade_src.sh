#!/bin/bash
# ncu --set full with source correlation for the shading kernels of a config-4 frame (second frame: launch-skip past the first)
mkdir -p gpurun_out
for k in gi_wf_shade_kernel rf_wf_gen_kernel rf_wf_shade_a_kernel shade_direct_kernel gi_wf_gen_kernel generate_gbuffer_kernel; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -f -o gpurun_out/r2_src_$k python tools/debug/one_frame.py config4_1080p_gi 2 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
