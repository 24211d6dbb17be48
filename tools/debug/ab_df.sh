for v in xy256_4 xy256_3 xy512_2 xy384_3 xy192_5 xy128_4; do
  export VXRT_CUDA_LIB=$PWD/voxeltracing_b200/libvxrt_cuda_$v.so
  echo "== $v"; python tools/debug/time_df.py 2>&1 | grep -E "stage=1 sx=0|stage=0 sx=0|stage=1 sx=1 sy=4|bit-exact again"
done
