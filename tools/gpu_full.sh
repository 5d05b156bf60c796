#!/bin/bash
# Full GPU check: smoke, every -m gpu test, default bench (with e2e + CPU baseline), reference arm, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" > gpurun_out/check.log
timeout -s KILL 300 python __graft_entry__.py smoke >> gpurun_out/check.log 2>&1; echo "smoke rc=$?" >> gpurun_out/check.log
echo "== gpu tests" >> gpurun_out/check.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider >> gpurun_out/check.log 2>&1; echo "tests rc=$?" >> gpurun_out/check.log
tail -15 gpurun_out/check.log
timeout -s KILL 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.json
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
