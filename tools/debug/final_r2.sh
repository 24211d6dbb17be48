#!/bin/bash
# round-2 closing run on one GPU: full GPU suite, smoke(), the default bench line and the CPU arm
out=gpurun_out; mkdir -p $out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > $out/r2_final_gpu_tests.log; cat $out/r2_final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/r2_final_smoke.log
python bench.py > $out/r2_final_bench.json 2> $out/r2_final_bench.err; tail -2 $out/r2_final_bench.err
python bench.py --impl reference > $out/r2_final_bench_ref.json 2> $out/r2_final_bench_ref.err; tail -2 $out/r2_final_bench_ref.err
python tools/debug/show_line.py $out/r2_final_bench.json
python - <<PY
import json
a=json.load(open("$out/r2_final_bench.json")); b=json.load(open("$out/r2_final_bench_ref.json"))
print("same config:", a["config"]==b["config"], "| ref value", round(b["value"],2), b["cpu_baseline"]["cores"], "cores | e2e ratio", round(a["e2e"]["value"]/b["value"],1), "| steps/warmup", a["steps"], a["warmup"])
PY
