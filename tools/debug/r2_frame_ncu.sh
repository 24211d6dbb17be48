#!/bin/bash
# one config-4 frame under ncu: per-launch duration, DRAM bytes, issue activity, lanes
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
tag=${1:-w}
shift
wl=${WORKLOAD:-config4_1080p_gi}
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_${tag}_frame_ncu.csv python tools/debug/one_frame.py $wl 2 "$@" > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_${tag}_frame_ncu.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki]), {})[r[mi]]=float(r[vi].replace(",",""))
items=sorted(d.items())
# the last frame = launches after the last initial_trace_kernel
last=max(i for (i,k),m in items if "initial_trace" in k)
tot=0; rd=0; wr=0
out=[]
for (i,k),m in items:
    if i<last: continue
    name=k.split("(")[0].replace("void ","").replace("<unnamed>::","")
    out.append("%3d %-44s %8.1f us  R %7.1f MB  W %7.1f MB  inst %6.1f M  lanes %5.2f  issue %5.1f %%  warps %5.1f %%  regs %3d  L1 %5.1f %%  L2 %5.1f %%" % (i, name[:44], m["gpu__time_duration.sum"]/1e3, m["dram__bytes_read.sum"]/1e6, m["dram__bytes_write.sum"]/1e6, m["sm__inst_executed.sum"]/1e6, m["smsp__thread_inst_executed_per_inst_executed.ratio"], m["smsp__issue_active.avg.pct_of_peak_sustained_active"], m["sm__warps_active.avg.pct_of_peak_sustained_active"], m["launch__registers_per_thread"], m["l1tex__t_sector_hit_rate.pct"], m["lts__t_sector_hit_rate.pct"]))
    tot+=m["gpu__time_duration.sum"]/1e3; rd+=m["dram__bytes_read.sum"]/1e6; wr+=m["dram__bytes_write.sum"]/1e6
out.append("frame: %.1f us under ncu (cold caches, serialised), DRAM read %.1f MB, written %.1f MB" % (tot, rd, wr))
open("gpurun_out/r2_${tag}_frame_ncu.txt","w").write("\n".join(out)+"\n")
print("\n".join(out))
PY
