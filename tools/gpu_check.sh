#!/bin/bash
# Run on the GPU box via gpurun: first-light + parity tests, each under its own hard timeout so that a
# hung kernel cannot hold the box until gpurun's limit.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" > gpurun_out/check.log
timeout -s KILL 300 python __graft_entry__.py smoke >> gpurun_out/check.log 2>&1; echo "smoke rc=$?" >> gpurun_out/check.log
echo "== 1d" >> gpurun_out/check.log
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "1d" --timeout 600 --timeout-method=thread -p no:cacheprovider >> gpurun_out/check.log 2>&1; echo "1d rc=$?" >> gpurun_out/check.log
echo "== rest" >> gpurun_out/check.log
timeout -s KILL 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not 1d" --timeout 900 --timeout-method=thread -p no:cacheprovider >> gpurun_out/check.log 2>&1; echo "rest rc=$?" >> gpurun_out/check.log
tail -60 gpurun_out/check.log
