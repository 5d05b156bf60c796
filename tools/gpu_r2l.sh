#!/bin/bash
# full -m gpu suite + the bench lines of every workload (new cpu_baseline / e2e legs)
mkdir -p gpurun_out
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2l_tests.log
for w in c2e c3 c4 c5a c5b; do
  timeout -s KILL 900 python bench.py --workload $w --steps 5 > gpurun_out/r2l_$w.json 2> gpurun_out/r2l_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/r2l_$w.err
  python - <<P
import json
d=json.load(open('gpurun_out/r2l_$w.json'))
r=d['roofline'] or {}
print('$w', round(d['value'],2), round(d['ms_per_step'],3), 'frac', r.get('frac'), 'fdram', r.get('frac_dram'), {k:round(v['ms_per_launch'],3) for k,v in (r.get('kernels') or {}).items()})
print('   cpu', d.get('cpu_baseline'))
print('   e2e', d.get('e2e'))
print('   parity', d.get('parity'))
P
done
