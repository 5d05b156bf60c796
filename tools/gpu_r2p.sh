#!/bin/bash
# per-element kappa split kernels with double-buffered rows: parity + c2e timing
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q --timeout 600 -p no:cacheprovider -k "per_element or unaligned or 1d" 2>&1 | tail -3
timeout -s KILL 600 python bench.py --workload c2e --steps 5 2>gpurun_out/r2p.err | tee gpurun_out/r2p_c2e.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2e', round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()}, d['parity'])"
tail -3 gpurun_out/r2p.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_c2e.csv python bench.py --workload c2e --steps 1 --warmup 3 --no-parity > gpurun_out/ncu_l2e.log 2>&1
python - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_c2e.csv')))
st=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
h=rows[st]
tot=collections.defaultdict(float);cnt=collections.Counter()
for r in rows[st+1:]:
    d=dict(zip(h,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(d['Metric Value'].replace(',',''))*{'ns':1e-3,'us':1,'ms':1e3}[d['Metric Unit']]
    tot[d['Kernel Name'][:60]]+=v;cnt[d['Kernel Name'][:60]]+=1
for k,v in sorted(tot.items(),key=lambda x:-x[1])[:8]: print('%9.1f us avg  n=%d  %s'%(v/cnt[k],cnt[k],k))
P
