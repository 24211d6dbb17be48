#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__t_sector_hit_rate.pct
for opt in trace_spill=0 trace_spill=8 trace_spill=0x080c; do
  ncu --metrics $M --clock-control none -k regex:"trace" --csv --log-file gpurun_out/r2_v_ncu_$opt.csv python tools/debug/one_frame.py config4_1080p_gi 2 $opt > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_v_ncu_$opt.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:70]), {})[r[mi]]=r[vi]
print("== $opt")
for (i,k),m in sorted(d.items()):
    print(i, k.split("(")[0][-60:], " ".join(f"{a.split('.')[0].split('__')[-1][:14]}={b}" for a,b in m.items()))
PY
done
