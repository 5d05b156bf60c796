#!/bin/bash
# A/B of the launch mode / poison kernel on config 2, the misfit sweep with the new geometry, and the failed test.
mkdir -p gpurun_out
B="python bench.py --steps 10 --no-e2e --no-cpu --no-parity --no-sweep"
for i in 1 2; do
$B > gpurun_out/ab_default_$i.json 2>/dev/null
DFE_PIPE_PLAIN_LAUNCH=1 $B > gpurun_out/ab_plain_$i.json 2>/dev/null
DFE_PIPE_NO_POISON=1 $B > gpurun_out/ab_nopoison_$i.json 2>/dev/null
DFE_PIPE_PLAIN_LAUNCH=1 DFE_PIPE_NO_POISON=1 $B > gpurun_out/ab_both_$i.json 2>/dev/null
done
for f in gpurun_out/ab_*.json; do echo -n "$f "; python -c "
import json,sys
d=json.load(open('$f')); k=d['roofline']['kernels']
print(round(d['ms_per_step'],3), round(k['solve1d_fwd']['ms_per_launch'],3), round(k['solve1d_bwd']['ms_per_launch'],3))"; done
python bench.py --workload c5a --steps 5 > gpurun_out/bench_c5a.json 2> gpurun_out/bench_c5a.err
python -c "
import json
d=json.load(open('gpurun_out/bench_c5a.json')); print(d['ms_per_step'], {k:v['ms_per_launch'] for k,v in d['roofline']['kernels'].items()})"
timeout -s KILL 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "large_chain or misfit" --timeout 900 -p no:cacheprovider 2>&1 | tail -15
