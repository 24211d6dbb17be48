#!/bin/bash
# adaptive hand-over of the queue trace kernels: bit identity, then pass times per threshold schedule and per build variant
mkdir -p gpurun_out
python -m pytest tests/test_gpu_shade.py -x -q -k "hand_over or capped" 2>&1 | tail -3 > gpurun_out/r2_u_spill_tests.log; cat gpurun_out/r2_u_spill_tests.log
export VXRT_SWEEP_OPTION=trace_spill
for v in "" spillocc10 spillocc12 spillch1 spillch4; do
  if [ -z "$v" ]; then unset VXRT_CUDA_LIB; tag=base; else export VXRT_CUDA_LIB=$PWD/voxeltracing_b200/libvxrt_cuda_$v.so; tag=$v; fi
  echo "== $tag" | tee -a gpurun_out/r2_u_spill_sweep.txt
  python tools/debug/sweep_caps.py config4_1080p_gi 0,4,8,12,16,8-8,12-8,16-8,16-12-8 2>&1 | tail -9 | tee -a gpurun_out/r2_u_spill_sweep.txt
done
