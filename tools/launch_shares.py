import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0].replace('<unnamed>::','').replace('void ','')
    ns=float(r[-1]); a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=ns
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:50s} n={v[0]:4d} avg={v[1]/v[0]/1e3:9.1f}us share={v[1]/tot*100:5.1f}%")
