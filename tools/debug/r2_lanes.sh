#!/bin/bash
# lane 2 / deferred reflection GI / lane copies: the tests that cover them, then the A / B timings
out=gpurun_out; mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_bench_configs.py tests/test_gpu_shade.py tests/test_gpu_trace.py -x -q -m gpu 2>&1 | tail -15 > $out/r2_zm_tests.log; cat $out/r2_zm_tests.log
timeout 200 python tools/debug/ab_lanes.py \
  lane2_direct=1,refl_defer_gi=1,copy_lanes=1 \
  lane2_direct=0,refl_defer_gi=0,copy_lanes=0 \
  lane2_direct=1,refl_defer_gi=0,copy_lanes=0 \
  lane2_direct=0,refl_defer_gi=1,copy_lanes=0 \
  lane2_direct=1,refl_defer_gi=1,copy_lanes=0 \
  lane2_direct=0,refl_defer_gi=0,copy_lanes=1 \
  lane2_direct=1,refl_defer_gi=1,copy_lanes=1,copies=0 \
  lane2_direct=1,refl_defer_gi=1,copy_lanes=1 2>&1 | tail -12 | tee $out/r2_zm_ab_lanes.txt
