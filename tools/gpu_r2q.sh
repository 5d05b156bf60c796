#!/bin/bash
# Chebyshev smoothing weights in the multigrid V-cycle: parity + iteration counts / timings, with and without
mkdir -p gpurun_out
for cheb in 1 0; do
  export DFE_MG_CHEB=$cheb
  for w in c4 c3; do
    for nu in 2 3; do
    DFE_MG_NU=$nu timeout -s KILL 600 python bench.py --workload $w --steps 3 --no-cpu --no-e2e 2>gpurun_out/r2q.err | tee gpurun_out/r2q_${w}_cheb${cheb}_nu${nu}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w cheb=$cheb nu=$nu', round(d['ms_per_step'],2), r['iterations'], round(r['us_per_iteration'],1))"
    done
  done
done
unset DFE_MG_CHEB
timeout -s KILL 1500 python -m pytest tests/test_gpu_mg.py -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -12
