#!/bin/bash
# usage: r2_tiles.sh N tag — config 5 strong scaling at N GPUs: column bands (1 per rank), 2 column bands per rank, row strips (4 per rank)
N=$1; tag=${2:-y}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
for mode in "cols 1" "cols 2" "rows 4" "rows 1"; do
  set -- $mode
  f=gpurun_out/r2_${tag}_n${N}_config5_$1$2
  $TR bench.py --gpus $N --steps 6 --warmup 3 --workload config5_4k_gi4 --no-svgf --no-cpu-baseline --tile-shape $1 --strips $2 > $f.json 2> $f.err; tail -2 $f.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_${tag}_n${N}_config5_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "no line:", e); continue
    print(f.split("/")[-1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["scaling"],
          "parity", d.get("parity_check"), "| gather", d.get("gather_check"))
    print("   pass_ms", {k: round(v, 3) for k, v in d["pass_ms"].items()}, "last", d.get("pass_ms_last_rank") and {k: round(v, 3) for k, v in d["pass_ms_last_rank"].items()})
PY
