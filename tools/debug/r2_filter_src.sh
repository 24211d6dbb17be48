#!/bin/bash
# ncu --set full with source correlation for the four most expensive screen-space filter kernels (config-4 frames at 1080p)
mkdir -p gpurun_out
for k in svgf_variance_kernel svgf_spatial_kernel shadow_filter_kernel reflection_denoise_kernel; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_src_$k python tools/debug/time_filters.py 0 > /dev/null 2>&1
done
ls -la gpurun_out/r2_src_*filter*.ncu-rep gpurun_out/r2_src_svgf*.ncu-rep gpurun_out/r2_src_reflection*.ncu-rep
