#!/bin/bash
# ncu source-level captures: config 5b rhs / gradient kernels; assembly tile kernel at 4 CTAs per SM
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "assembly" --timeout 600 -p no:cacheprovider 2>&1 | tail -2
timeout -s KILL 600 python bench.py --workload c4 --steps 5 --no-cpu --no-e2e 2>gpurun_out/bench_c4.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c4', round(d['ms_per_step'],2), r['iterations'], {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
SHORT4="python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tile -s 2 -c 1 -f -o gpurun_out/prof_assemble_tile4 $SHORT4 > gpurun_out/ncu_assemble_tile.log 2>&1; echo "ncu assemble rc=$?"
SHORT5="python bench.py --workload c5b --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_band_rhs_fwd2|k_band_grad2" -s 2 -c 2 -f -o gpurun_out/prof_band2 $SHORT5 > gpurun_out/ncu_band2.log 2>&1; echo "ncu band2 rc=$?"
