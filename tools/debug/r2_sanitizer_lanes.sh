#!/bin/bash
# compute-sanitizer memcheck over the lane test of the final build (lane 2, deferred reflection GI kernels, per-lane copy streams), refl_spp = 1
out=gpurun_out; mkdir -p $out
timeout 200 compute-sanitizer --tool memcheck --log-file $out/r2_memcheck_lanes.raw python -m pytest -q -x \
  "tests/test_gpu_bench_configs.py::test_lane2_direct_deferred_reflection_gi_and_lane_copies_do_not_change_a_bit[1]" 2>&1 | tail -3 > $out/r2_memcheck_lanes.log
tail -2 $out/r2_memcheck_lanes.raw >> $out/r2_memcheck_lanes.log
cat $out/r2_memcheck_lanes.log
