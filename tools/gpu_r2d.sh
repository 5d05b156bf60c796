#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "assembly or 2d or pcg" --timeout 600 -p no:cacheprovider 2>&1 | tail -8
timeout -s KILL 1500 python -m pytest tests/test_gpu_mg.py -m gpu -q -x --timeout 900 -p no:cacheprovider --durations=5 2>&1 | tail -25
for nu in 2 1 3; do
DFE_MG_NU=$nu timeout -s KILL 600 python bench.py --workload c4 --steps 3 --no-cpu --no-e2e 2>gpurun_out/bench_c4.err | tee gpurun_out/bench_c4_nu$nu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c4 nu=$nu', round(d['ms_per_step'],2), r['iterations'], round(r['us_per_iteration'],1), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
done
tail -3 gpurun_out/bench_c4.err
timeout -s KILL 600 python bench.py --workload c3 --steps 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c3', round(d['ms_per_step'],2), r['iterations'], round(r['us_per_iteration'],1), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()})"
DFE_SOLVER2D=jacobi timeout -s KILL 600 python bench.py --workload c3 --steps 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c3 jacobi', round(d['ms_per_step'],2), r['iterations'], round(r['us_per_iteration'],1))"
# ncu: assembly kernel and the MG kernel
SHORT4="python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_grid -s 2 -c 1 -f -o gpurun_out/prof_assemble $SHORT4 > gpurun_out/ncu_assemble.log 2>&1; echo "ncu assemble rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_mgpcg -s 2 -c 1 -f -o gpurun_out/prof_mgpcg $SHORT4 > gpurun_out/ncu_mgpcg.log 2>&1; echo "ncu mgpcg rc=$?"
