#!/bin/bash
# element-parallel structured assembly (k_assemble_tile) + config 5b kernels: parity, A/B timings, ncu of the assembly
mkdir -p gpurun_out
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mg.py -m gpu -q -x -k "batch or band or assembly or error or 2d or mg" --timeout 600 -p no:cacheprovider 2>&1 | tail -6
timeout -s KILL 600 python bench.py --workload c5b --steps 5 --no-cpu --no-e2e 2>gpurun_out/bench_c5b.err | tee gpurun_out/r2i_c5b.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5b', round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
for mode in tile rows; do
  if [ $mode = rows ]; then export DFE_ASSEMBLE_ROWS=1; else unset DFE_ASSEMBLE_ROWS; fi
  timeout -s KILL 600 python bench.py --workload c4 --steps 5 --no-cpu --no-e2e 2>gpurun_out/bench_c4.err | tee gpurun_out/r2i_c4_$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c4 $mode', round(d['ms_per_step'],2), r['iterations'], {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
done
unset DFE_ASSEMBLE_ROWS
SHORT4="python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tile -s 2 -c 1 -f -o gpurun_out/prof_assemble_tile $SHORT4 > gpurun_out/ncu_assemble_tile.log 2>&1; echo "ncu assemble rc=$?"
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
for k,v in sorted(tot.items(),key=lambda x:-x[1])[:8]: print('%9.1f us avg  n=%d  %s'%(v/cnt[k],cnt[k],k))
P
