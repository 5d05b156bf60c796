#!/bin/bash
# single fence after the mbarrier inits: c5b / c2 / c2e timings + pipeline parity
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -p no:cacheprovider 2>&1 | tail -2
for w in c5b c2e c2; do
  timeout -s KILL 600 python bench.py --workload $w --steps 5 --no-cpu --no-e2e --no-sweep 2>gpurun_out/r2m.err | tee gpurun_out/r2m_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w', round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
done
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5b.csv python bench.py --workload c5b --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_l5b.log 2>&1
grep -o "k_band_solve_mma[^\"]*\"[^\"]*\"[^\"]*\"[^\"]*\"[^\"]*\"[^\"]*\"[0-9.,]*" gpurun_out/launches_c5b.csv | tail -2
