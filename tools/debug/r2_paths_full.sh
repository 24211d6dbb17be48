#!/bin/bash
# ncu --set full of the dominant kernels of the final build (second frame of one_frame.py): both launches of the GI path-ray kernel and the
# reflection closest-hit kernel
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:'wf_trace_paths_kernel|rf_wf_trace_kernel' -s 3 -c 3 -f -o gpurun_out/r2_zr_trace_full python tools/debug/one_frame.py config4_1080p_gi 2 > /dev/null 2>&1
ncu -i gpurun_out/r2_zr_trace_full.ncu-rep --page raw --csv > gpurun_out/r2_zr_trace_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_zr_trace_full_raw.csv > gpurun_out/r2_zr_trace_ncu_full_summary.txt
grep -E "Kernel Name|gpu__time_duration|dram__bytes|issue_active|thread_inst_executed_per" gpurun_out/r2_zr_trace_ncu_full_summary.txt | cut -c1-160
