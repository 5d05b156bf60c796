#!/bin/bash
# config 5b: register-resident rhs / gradient kernels around the band solve — parity, then new vs old (DFE_BAND_OLD=1)
mkdir -p gpurun_out
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch or band or assembly or error" --timeout 600 -p no:cacheprovider 2>&1 | tail -6
DFE_BAND_REG=1 timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch or band" --timeout 600 -p no:cacheprovider 2>&1 | tail -2
for mode in new reg old; do
  unset DFE_BAND_OLD DFE_BAND_REG
  if [ $mode = old ]; then export DFE_BAND_OLD=1; fi
  if [ $mode = reg ]; then export DFE_BAND_REG=1; fi
  timeout -s KILL 600 python bench.py --workload c5b --steps 5 --no-cpu --no-e2e 2>gpurun_out/bench_c5b.err | tee gpurun_out/r2h_c5b_$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5b $mode', round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
done
unset DFE_BAND_OLD DFE_BAND_REG
tail -3 gpurun_out/bench_c5b.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5b.csv python bench.py --workload c5b --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_l5b.log 2>&1; echo "launches c5b rc=$?"
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
    tot[d['Kernel Name'][:60]]+=v;cnt[d['Kernel Name'][:60]]+=1
for k,v in sorted(tot.items(),key=lambda x:-x[1])[:12]: print('%9.1f us avg  n=%d  %s'%(v/cnt[k],cnt[k],k))
P
