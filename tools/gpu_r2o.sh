#!/bin/bash
# blocked banded Cholesky: parity + c5b timings (blocked vs row-by-row) + heat tests
mkdir -p gpurun_out
timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout -s KILL 900 python -m pytest tests/test_gpu_heat.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -p no:cacheprovider -k "heat or band or batch or error" 2>&1 | tail -3
for mode in blk nofuse rows; do
  unset DFE_BAND_FACTOR_ROWS DFE_BAND_NOFUSE
  if [ $mode = rows ]; then export DFE_BAND_FACTOR_ROWS=1; fi
  if [ $mode = nofuse ]; then export DFE_BAND_NOFUSE=1; fi
  timeout -s KILL 600 python bench.py --workload c5b --steps 10 --no-cpu --no-e2e 2>gpurun_out/r2o.err | tee gpurun_out/r2o_c5b_$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5b $mode', round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
done
unset DFE_BAND_FACTOR_ROWS DFE_BAND_NOFUSE
tail -3 gpurun_out/r2o.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5b.csv python bench.py --workload c5b --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_l5b.log 2>&1
grep "k_band_factor" gpurun_out/launches_c5b.csv | tail -2 | cut -c1-60,200-400
