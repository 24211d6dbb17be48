for cfg in "default" "NCCL_MAX_CTAS=8" "NCCL_MAX_CTAS=16" "NCCL_MAX_CTAS=4"; do
  if [ "$cfg" = "default" ]; then unset NCCL_MAX_CTAS; else export $cfg; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$cfg', 'value=%.0f'%d['value'], 'ms/step=%.3f'%d['ms_per_step'], 'e2e=%.0f'%d['e2e']['value'], 'numa', d['e2e'].get('rank0_numa_node'))"
done
