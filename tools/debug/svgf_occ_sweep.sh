python -m pytest tests/test_gpu_svgf.py -q 2>&1 | tail -30 > gpurun_out/svgf_tests.log; tail -c 3000 gpurun_out/svgf_tests.log
for v in "" svgf_occ5 svgf_occ6; do
  if [ -n "$v" ]; then export VXRT_CUDA_LIB=$PWD/voxeltracing_b200/libvxrt_cuda_$v.so; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_occ_$v.json 2> gpurun_out/bench_occ_$v.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_occ_$v.json').read())['svgf']; print('$v', d['ms_per_chain'], {k: round(s['ms_per_launch'],4) for k,s in d['stages'].items()})"
done
