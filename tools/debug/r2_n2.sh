#!/bin/bash
# 2-GPU checks: the NCCL sharding test, then bench lines (frames / push, frames / nccl, config 5 tiles)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharding.py -x -q 2>&1 | tail -4 > gpurun_out/r2_m_sharding_tests.log
cat gpurun_out/r2_m_sharding_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_m_n2_push.json 2> gpurun_out/r2_m_n2_push.err; tail -3 gpurun_out/r2_m_n2_push.err
$TR bench.py --gpus 2 --steps 10 --warmup 3 --gather nccl --outputs all --no-parity-check > gpurun_out/r2_m_n2_nccl_all.json 2> gpurun_out/r2_m_n2_nccl.err; tail -3 gpurun_out/r2_m_n2_nccl.err
$TR bench.py --gpus 2 --steps 4 --warmup 3 --workload config5_4k_gi4 > gpurun_out/r2_m_n2_config5_tiles.json 2> gpurun_out/r2_m_n2_config5.err; tail -3 gpurun_out/r2_m_n2_config5.err
python - <<'PY'
import json
for f in ("r2_m_n2_push", "r2_m_n2_nccl_all", "r2_m_n2_config5_tiles"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "no line:", e); continue
    print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "scaling", d["scaling"],
          "parity", d.get("parity_check"), "gather", d.get("gather_check"))
    print("   pass_ms", {k: round(v, 3) for k, v in d["pass_ms"].items()}, "last rank", d.get("pass_ms_last_rank") and {k: round(v, 3) for k, v in d["pass_ms_last_rank"].items()})
    print("   df_sharded", d.get("df_regen_sharded"))
PY
