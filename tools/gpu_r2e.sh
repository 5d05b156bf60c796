#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batched or assembly or error" --timeout 600 -p no:cacheprovider 2>&1 | tail -6
for mode in mma scalar; do
  if [ $mode = scalar ]; then export DFE_BAND_SCALAR=1; else unset DFE_BAND_SCALAR; fi
  timeout -s KILL 600 python bench.py --workload c5b --steps 5 --no-cpu --no-e2e 2>gpurun_out/bench_c5b.err | tee gpurun_out/bench_c5b_$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5b $mode', round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
done
unset DFE_BAND_SCALAR
tail -3 gpurun_out/bench_c5b.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5b.csv python bench.py --workload c5b --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_l5b.log 2>&1; echo "launches c5b rc=$?"
for w in c3 c4; do
timeout -s KILL 600 python bench.py --workload $w --steps 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w', round(d['ms_per_step'],2), r['iterations'], round(r['us_per_iteration'],1), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
done
DFE_SOLVER2D=jacobi timeout -s KILL 600 python bench.py --workload c3 --steps 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c3 jacobi', round(d['ms_per_step'],2), r['iterations'], round(r['us_per_iteration'],1))"
