#!/bin/bash
# no-output adjoint variants (ring slots released after phase B): parity, then config 5a / config 2 per DFE_PIPE_CFG
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/r2f_tests.log 2>&1; echo "pipeline tests rc=$?"
tail -5 gpurun_out/r2f_tests.log
for cfg in 0 1 2; do
  DFE_PIPE_CFG=$cfg timeout -s KILL 300 python bench.py --workload c5a --steps 10 --warmup 3 > gpurun_out/r2f_c5a_$cfg.json 2> gpurun_out/r2f_c5a_$cfg.err; echo "c5a cfg$cfg rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/r2f_c5a_$cfg.json'))
print('c5a cfg$cfg', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})
P
done
for cfg in 0 1 2; do
  DFE_PIPE_CFG=$cfg timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-sweep --no-parity > gpurun_out/r2f_c2_$cfg.json 2> gpurun_out/r2f_c2_$cfg.err; echo "c2 cfg$cfg rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/r2f_c2_$cfg.json'))
print('c2 cfg$cfg', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()}, 'e2e', round(d['e2e']['value']), 'full', round(d['e2e_full']['value']))
P
done
