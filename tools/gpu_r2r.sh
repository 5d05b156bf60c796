#!/bin/bash
# split gradient kernel (GK / GF halves) vs fused: parity + c5b timing
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_heat.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -p no:cacheprovider -k "heat or band or batch or error" 2>&1 | tail -2
for mode in split fused; do
  unset DFE_BAND_GRAD_FUSED
  if [ $mode = fused ]; then export DFE_BAND_GRAD_FUSED=1; fi
  timeout -s KILL 600 python bench.py --workload c5b --steps 10 --no-cpu --no-e2e 2>gpurun_out/r2r.err | tee gpurun_out/r2r_c5b_$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5b $mode', round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
done
unset DFE_BAND_GRAD_FUSED
tail -3 gpurun_out/r2r.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5b.csv python bench.py --workload c5b --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_l5b.log 2>&1
python - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_c5b.csv')))
st=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
h=rows[st]
tot=collections.defaultdict(float);cnt=collections.Counter()
for r in rows[st+1:]:
    d=dict(zip(h,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(d['Metric Value'].replace(',',''))*{'ns':1e-3,'us':1,'ms':1e3}[d['Metric Unit']]
    tot[d['Kernel Name'][:70]]+=v;cnt[d['Kernel Name'][:70]]+=1
for k,v in sorted(tot.items(),key=lambda x:-x[1])[:9]: print('%9.1f us avg  n=%d  %s'%(v/cnt[k],cnt[k],k))
P
