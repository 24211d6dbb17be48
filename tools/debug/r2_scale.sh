#!/bin/bash
# usage: r2_scale.sh N  — bench lines at N GPUs: config 4 frames (push), config 4 frames (nccl, all outputs: round 1's path), config 5 tiles
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; fi
$TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_t_n${N}_config4.json 2> gpurun_out/r2_t_n${N}_config4.err; tail -2 gpurun_out/r2_t_n${N}_config4.err
if [ "$N" != "1" ]; then
  $TR bench.py --gpus $N --steps 20 --warmup 3 --gather nccl --outputs all --no-parity-check --no-cpu-baseline > gpurun_out/r2_t_n${N}_config4_nccl_all.json 2> gpurun_out/r2_t_n${N}_nccl.err; tail -2 gpurun_out/r2_t_n${N}_nccl.err
  python -m pytest tests/test_gpu_sharding.py -x -q 2>&1 | tail -3 > gpurun_out/r2_t_n${N}_sharding_tests.log
fi
$TR bench.py --gpus $N --steps 6 --warmup 3 --workload config5_4k_gi4 --no-svgf > gpurun_out/r2_t_n${N}_config5.json 2> gpurun_out/r2_t_n${N}_config5.err; tail -2 gpurun_out/r2_t_n${N}_config5.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_t_n${N}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "no line:", e); continue
    print(f.split("/")[-1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["scaling"],
          "parity", d.get("parity_check"), "| gather", d.get("gather_check"), "| d2h GB/s/rank", round(d["e2e"].get("d2h_gbs_per_rank", 0), 1))
    print("   pass_ms", {k: round(v, 3) for k, v in d["pass_ms"].items()}, "last", d.get("pass_ms_last_rank") and {k: round(v, 3) for k, v in d["pass_ms_last_rank"].items()})
    if d.get("df_regen_sharded"): print("   df_sharded us", round(d["df_regen_sharded"]["us_per_regeneration"], 1), {k: round(v, 1) for k, v in d["df_regen_sharded"]["phases_us"].items()})
PY
