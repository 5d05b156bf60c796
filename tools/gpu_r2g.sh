#!/bin/bash
# forward kernel with one ring slot less (PF = 0) and a larger chunk: config 2 / 5a per DFE_PIPE_CFG, parity first
mkdir -p gpurun_out
for cfg in 5 6; do
DFE_PIPE_CFG=$cfg timeout -s KILL 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q --timeout 600 -p no:cacheprovider -k "many_iterations or misfit" > gpurun_out/r2g_tests_$cfg.log 2>&1; echo "pipeline tests cfg$cfg rc=$?"
tail -2 gpurun_out/r2g_tests_$cfg.log
done
for cfg in 0 5 6; do
  DFE_PIPE_CFG=$cfg timeout -s KILL 300 python bench.py --workload c5a --steps 10 --warmup 3 > gpurun_out/r2g_c5a_$cfg.json 2> gpurun_out/r2g_c5a_$cfg.err; echo "c5a cfg$cfg rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/r2g_c5a_$cfg.json'))
print('c5a cfg$cfg', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})
P
  DFE_PIPE_CFG=$cfg timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-sweep --no-e2e > gpurun_out/r2g_c2_$cfg.json 2> gpurun_out/r2g_c2_$cfg.err; echo "c2 cfg$cfg rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/r2g_c2_$cfg.json'))
print('c2 cfg$cfg', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()}, d['parity']['max_rel'])
P
done
