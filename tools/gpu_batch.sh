#!/bin/bash
# Batched small-system kernel: 2-D parity tests, then the config-5b bench line.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "2d or pcg or general" --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/check2d.log 2>&1; echo "2d tests rc=$?"; tail -25 gpurun_out/check2d.log
timeout -s KILL 600 python bench.py --workload c5b --steps 3 > gpurun_out/bench_c5b.json 2> gpurun_out/bench_c5b.err; echo "c5b rc=$?"; tail -c 1500 gpurun_out/bench_c5b.json; tail -5 gpurun_out/bench_c5b.err
