#!/bin/bash
# round-1 closing run on the GPU box: full GPU suite, compute-sanitizer over the kernels added late in the round, ncu captures, bench lines.
# usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/debug/final_r1.sh f2'
tag=${1:-f2}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/gpu_tests_$tag.log 2>&1; tail -3 $out/gpu_tests_$tag.log
timeout 400 compute-sanitizer --tool memcheck --log-file $out/memcheck_filters_$tag.raw python -m pytest -q -x tests/test_gpu_lpv.py tests/test_gpu_world.py \
    "tests/test_gpu_svgf.py::test_prespatial_pass" tests/test_gpu_shadow_filter.py tests/test_gpu_refl_filter.py > $out/memcheck_filters_$tag.log 2>&1
tail -2 $out/memcheck_filters_$tag.log; tail -1 $out/memcheck_filters_$tag.raw
timeout 300 compute-sanitizer --tool racecheck --log-file $out/racecheck_lpv_$tag.raw python -m pytest -q -x tests/test_gpu_lpv.py::test_repropagate_matches_golden_and_oracle \
    tests/test_gpu_lpv.py::test_repropagate_other_dims "tests/test_gpu_svgf.py::test_prespatial_pass" > $out/racecheck_lpv_$tag.log 2>&1
tail -2 $out/racecheck_lpv_$tag.log; tail -1 $out/racecheck_lpv_$tag.raw
cat > /tmp/lpv_run.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, "tests")
import svgf_util as sv
from voxeltracing_b200 import abi, engine, host_api
c = engine.Context(0)
blocks = host_api.gen_world("rooms", 2)
rng = np.random.default_rng(4)
nz, ny, nx = blocks.shape
blocks[rng.integers(1, nz, 2000), rng.integers(1, ny, 2000), rng.integers(1, nx, 2000)] = 12
t = np.full((6, 128), -1, dtype=np.int32); t[3, 12] = 0
c.set_block_data(t); c.upload_world(blocks)
for limit in (4, 8):
    c.lpv_repropagate(None, limit)
lamp = np.argwhere(blocks == 12)[1000]
z, y, x = (int(v) for v in lamp)
c.edit_blocks(np.array([[x, y, z, 0]], dtype=np.int32)); c.lpv_edit(0, (x, y, z), 12, 8)
c.edit_blocks(np.array([[x, y, z, 12]], dtype=np.int32)); c.lpv_edit(1, (x, y, z), 12, 8)
# the 3 x 3 pre-pass at 1080p on synthetic attachments
W, H = 1920, 1080
g = np.random.default_rng(1)
c.write_attachment(abi.ATT_INITIAL_T, (10 + 5 * g.random((H, W))).astype(np.float16)); c.write_attachment(abi.ATT_INITIAL_NORMAL, (g.integers(0, 6, (H, W)) * 26).astype(np.uint8))
c.write_attachment(abi.ATT_INITIAL_BLOCK, np.full((H, W), 3, np.uint8))
c.write_set(abi.ATT_GI_SH, {"sh": g.random((H, W, 4)).astype(np.float16), "cocg": g.random((H, W, 2)).astype(np.float16), "x": g.random((H, W)).astype(np.float16),
                            "aosky": g.integers(0, 256, (H, W, 2)).astype(np.uint8)})
cam = host_api.camera([192.0, 75.0, 192.0], 30.0, -15.0, W / H)
p = sv.prespatial_params(cam); p.width, p.height = W, H
for _ in range(3):
    c.svgf_prespatial(p)
c.close()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lpv_|svgf_prespatial|lights_kernel|scan_counts' -c 80 -o $out/lpv_$tag -f python /tmp/lpv_run.py > $out/ncu_lpv_$tag.log 2>&1; tail -2 $out/ncu_lpv_$tag.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --profile > $out/bench_profile_$tag.log 2>&1
timeout 300 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -c 600 $out/bench_$tag.json
for w in config3_1080p_direct config5_4k_gi4; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-svgf > $out/bench_${tag}_$w.json 2> $out/bench_${tag}_$w.err; tail -c 300 $out/bench_${tag}_$w.json
done
