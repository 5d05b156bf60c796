#!/bin/bash
# Poll back-off sweeps: PCG all-reduce (configs 3 / 4) and the fold-warp polls of the pipelined 1-D kernel (config 2).
mkdir -p gpurun_out
for bo in ${PCG_BO:-100 250 500 1000}; do
  for w in c3 c4; do
    DFE_PCG_BACKOFF=$bo timeout -s KILL 300 python bench.py --workload $w --steps 2 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pcg backoff', $bo, '$w', 'us/it %.2f' % d['roofline']['us_per_iteration'])"
  done
done
for bo in ${PIPE_BO:-0 100 250 500}; do
  DFE_PIPE_BACKOFF=$bo timeout -s KILL 120 python bench.py --steps 5 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pipe backoff', $bo, 'solves/s %.0f' % d['value'], {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"
done
