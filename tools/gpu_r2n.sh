#!/bin/bash
# heat-equation stepping (dfe_band_solve), TRSM with skipped zero tiles, sweep with the fused Adam kernel
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_heat.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -p no:cacheprovider -k "heat or band or batch" 2>&1 | tail -3
timeout -s KILL 300 python examples/heat_2d.py 2>&1 | tail -5
for w in c5b c5a; do
  timeout -s KILL 600 python bench.py --workload $w --steps 10 --no-cpu --no-e2e 2>gpurun_out/r2n.err | tee gpurun_out/r2n_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w', round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()}, d['gpu_launches'])"
done
tail -3 gpurun_out/r2n.err
