#!/bin/bash
# usage: r2_final_scale.sh N [tag] — bench lines of the round's final build at N GPUs: config 4 (frames; push), config 5 (N > 1: one column band
# per rank, strong scaling), the NCCL sharding tests and the all-ranks device-to-host copy roof
N=$1; tag=${2:-zc}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; fi
$TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_${tag}_n${N}_config4.json 2> gpurun_out/r2_${tag}_n${N}_config4.err; tail -2 gpurun_out/r2_${tag}_n${N}_config4.err
$TR bench.py --gpus $N --steps 6 --warmup 3 --workload config5_4k_gi4 --no-svgf > gpurun_out/r2_${tag}_n${N}_config5.json 2> gpurun_out/r2_${tag}_n${N}_config5.err; tail -2 gpurun_out/r2_${tag}_n${N}_config5.err
if [ "$N" != "1" ]; then
  python -m pytest tests/test_gpu_sharding.py -x -q 2>&1 | tail -3 > gpurun_out/r2_${tag}_n${N}_sharding_tests.log; cat gpurun_out/r2_${tag}_n${N}_sharding_tests.log
  $TR tools/debug/pcie_bw_n.py > gpurun_out/r2_${tag}_n${N}_d2h_roof.json 2> /dev/null; cat gpurun_out/r2_${tag}_n${N}_d2h_roof.json
fi
python tools/debug/show_line.py gpurun_out/r2_${tag}_n${N}_config4.json gpurun_out/r2_${tag}_n${N}_config5.json
