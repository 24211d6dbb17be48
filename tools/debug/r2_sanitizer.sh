#!/bin/bash
# round 2: compute-sanitizer over what the round added — rectangle tiles, rect copies, the GI side-stream overlap, pass overlap, the hand-over
# and capped trace passes, the second-generation DF kernels (memcheck everywhere; racecheck on the DF kernels' shared-memory phases)
out=gpurun_out; mkdir -p $out
timeout 900 compute-sanitizer --tool memcheck --log-file $out/r2_memcheck_tiles_overlap.raw python -m pytest -q -x tests/test_gpu_trace.py::test_rectangle_tiles_equal_full_frame_and_rect_copy \
  "tests/test_gpu_shade.py::test_fused_final_gi_kernel_is_bit_identical" "tests/test_gpu_shade.py::test_adaptive_hand_over_does_not_change_a_bit" \
  "tests/test_gpu_shade.py::test_capped_trace_passes_do_not_change_a_bit" 2>&1 | tail -3 > $out/r2_memcheck_tiles_overlap.log
grep -c "ERROR SUMMARY: 0 errors" $out/r2_memcheck_tiles_overlap.raw >> $out/r2_memcheck_tiles_overlap.log; tail -2 $out/r2_memcheck_tiles_overlap.raw >> $out/r2_memcheck_tiles_overlap.log
timeout 600 compute-sanitizer --tool memcheck --log-file $out/r2_memcheck_df.raw python -m pytest -q -x tests/test_gpu_df.py -k "version or edit" 2>&1 | tail -3 > $out/r2_memcheck_df.log
tail -2 $out/r2_memcheck_df.raw >> $out/r2_memcheck_df.log
timeout 600 compute-sanitizer --tool racecheck --log-file $out/r2_racecheck_df.raw python -m pytest -q -x tests/test_gpu_df.py -k "version" 2>&1 | tail -3 > $out/r2_racecheck_df.log
tail -2 $out/r2_racecheck_df.raw >> $out/r2_racecheck_df.log
cat $out/r2_memcheck_tiles_overlap.log $out/r2_memcheck_df.log $out/r2_racecheck_df.log
