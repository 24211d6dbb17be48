for v in "" occ4 occ5 occ6; do
  if [ -z "$v" ]; then unset VXRT_CUDA_LIB; tag=base; else export VXRT_CUDA_LIB=$PWD/voxeltracing_b200/libvxrt_cuda_$v.so; tag=$v; fi
  echo "== $tag"; python tools/debug/sweep_wf.py 2>&1 | tail -1
done
