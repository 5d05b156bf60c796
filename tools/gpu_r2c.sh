#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "assembly or 2d" --timeout 600 -p no:cacheprovider 2>&1 | tail -8
timeout -s KILL 1500 python -m pytest tests/test_gpu_mg.py -m gpu -q -x --timeout 900 -p no:cacheprovider --durations=8 2>&1 | tail -60
timeout -s KILL 600 python bench.py --workload c4 --steps 3 --no-cpu --no-e2e > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
head -c 3000 gpurun_out/bench_c4.json; tail -5 gpurun_out/bench_c4.err
DFE_ASSEMBLE_GENERAL=1 timeout -s KILL 600 python bench.py --workload c4 --steps 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('general assemble:', d['roofline']['kernels'].get('assemble'))"
timeout -s KILL 600 python bench.py --workload c3 --steps 3 --no-cpu --no-e2e > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
head -c 2000 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
timeout -s KILL 600 python bench.py --steps 10 --no-e2e --no-cpu --no-parity --no-sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print('c2', round(d['ms_per_step'],3), round(k['solve1d_fwd']['ms_per_launch'],3), round(k['solve1d_bwd']['ms_per_launch'],3))"
