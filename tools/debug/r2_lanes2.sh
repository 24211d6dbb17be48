#!/bin/bash
# material G-buffer on lane 1: the tests that cover the lanes, then the A / B timings
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_bench_configs.py -x -q -m gpu 2>&1 | tail -15 > $out/r2_zo_tests.log; cat $out/r2_zo_tests.log
timeout 200 python tools/debug/ab_lanes.py \
  lane1_gbuffer=1 \
  lane1_gbuffer=0 \
  lane1_gbuffer=1,wf_bands=3 \
  lane1_gbuffer=1,wf_bands=1 \
  lane1_gbuffer=1,wf_bands=2 2>&1 | tail -12 | tee $out/r2_zo_ab_lanes.txt
