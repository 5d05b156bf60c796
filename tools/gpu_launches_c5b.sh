#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5b.csv python bench.py --workload c5b --steps 1 --warmup 1 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/launches_c5b.csv")))
st=next(i for i,r in enumerate(rows) if r and r[0]=="ID"); hdr=rows[st]
tot=collections.defaultdict(float); cnt=collections.Counter()
for r in rows[st+1:]:
    d=dict(zip(hdr,r))
    if d.get("Metric Name")!="gpu__time_duration.sum": continue
    v=float(d["Metric Value"].replace(",",""))*{"ns":1e-3,"us":1.0,"ms":1e3,"s":1e6}[d["Metric Unit"]]
    tot[d["Kernel Name"][:70]]+=v; cnt[d["Kernel Name"][:70]]+=1
for k,v in sorted(tot.items(),key=lambda x:-x[1])[:9]: print("%-72s n=%3d avg %9.1f us"%(k,cnt[k],v/cnt[k]))
PY
