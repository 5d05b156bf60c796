#!/bin/bash
# P2 extension + tile assembly kernel with the per-mesh geometry range flag: tests, example, assembly timing
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_p2.py -x -q -m gpu -k "assembly or p2 or 2d" 2>&1 | tail -3
timeout -s KILL 300 python examples/convergence_2d.py 2>&1 | tail -12
rm -f gpurun_out/r2v_asm.jsonl
for n in 1024 2048; do
  timeout -s KILL 300 python tools/asm_bench.py $n 30 2>>gpurun_out/r2v.err | tee -a gpurun_out/r2v_asm.jsonl
done
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_assemble --csv python tools/asm_bench.py 1024 2 2>/dev/null | grep k_assemble | tail -4
