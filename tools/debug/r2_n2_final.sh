#!/bin/bash
# N = 2 check of the final build: config 4 frames (push gather, parity + checksum checks) and config 5 column tiles
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-svgf > gpurun_out/r2_zq_n2_config4.json 2> gpurun_out/r2_zq_n2_config4.err; tail -2 gpurun_out/r2_zq_n2_config4.err
$TR bench.py --gpus 2 --steps 6 --warmup 3 --workload config5_4k_gi4 --no-svgf --no-cpu-baseline > gpurun_out/r2_zq_n2_config5.json 2> gpurun_out/r2_zq_n2_config5.err; tail -2 gpurun_out/r2_zq_n2_config5.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_zq_n2_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "no line:", e); continue
    print(f.split("/")[-1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["scaling"],
          "parity", d.get("parity_check"), "| gather", d.get("gather_check"), "| d2h GB/s/rank", round(d["e2e"].get("d2h_gbs_per_rank", 0), 1))
PY
