#!/bin/bash
# usage: r2_tiles1.sh N tag [extra bench args] — config 5 strong scaling at N GPUs, default tiling (one column band per rank)
N=$1; tag=${2:-y}; shift; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
f=gpurun_out/r2_${tag}_n${N}_config5_cols1
$TR bench.py --gpus $N --steps 6 --warmup 3 --workload config5_4k_gi4 --no-svgf --no-cpu-baseline "$@" > $f.json 2> $f.err; tail -2 $f.err
python - <<PY
import json
d = json.load(open("$f.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d["scaling"], "parity", d.get("parity_check"), "| gather", d.get("gather_check"))
print("   pass_ms", {k: round(v, 3) for k, v in d["pass_ms"].items()}, "last", d.get("pass_ms_last_rank") and {k: round(v, 3) for k, v in d["pass_ms_last_rank"].items()})
print("   step_ms", d.get("step_ms"))
PY
