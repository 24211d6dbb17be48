#!/bin/bash
# register allocation of the queue trace kernels: VX_TRACE_OCC variants (min resident CTAs asked of ptxas) against the default build
mkdir -p gpurun_out
for v in "" "$@"; do
  if [ -z "$v" ]; then unset VXRT_CUDA_LIB; tag=base; else export VXRT_CUDA_LIB=$PWD/voxeltracing_b200/libvxrt_cuda_$v.so; tag=$v; fi
  [ -n "$v" ] && [ ! -f "$VXRT_CUDA_LIB" ] && continue
  echo "== $tag" | tee -a gpurun_out/r2_x_tocc_sweep.txt
  python tools/debug/sweep_caps.py config4_1080p_gi 0,0 2>&1 | tail -2 | tee -a gpurun_out/r2_x_tocc_sweep.txt
done
