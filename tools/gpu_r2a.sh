#!/bin/bash
# Round 2, first GPU call: smoke, the new pipeline-regime tests, the old 1-D tests, default bench (short).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" > gpurun_out/check.log
timeout -s KILL 300 python __graft_entry__.py smoke >> gpurun_out/check.log 2>&1; echo "smoke rc=$?" >> gpurun_out/check.log
echo "== pipeline tests" >> gpurun_out/check.log
timeout -s KILL 1500 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider --durations=10 >> gpurun_out/check.log 2>&1; echo "pipeline rc=$?" >> gpurun_out/check.log
echo "== parity tests" >> gpurun_out/check.log
timeout -s KILL 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider --durations=10 >> gpurun_out/check.log 2>&1; echo "parity rc=$?" >> gpurun_out/check.log
timeout -s KILL 900 python bench.py --steps 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --workload c5a --steps 5 > gpurun_out/bench_c5a.json 2> gpurun_out/bench_c5a.err; echo "c5a rc=$?"
grep -v "^$" gpurun_out/check.log | tail -80
head -c 6000 gpurun_out/bench.json; echo; tail -5 gpurun_out/bench.err
head -c 3000 gpurun_out/bench_c5a.json; echo; tail -5 gpurun_out/bench_c5a.err
