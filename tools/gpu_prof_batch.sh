#!/bin/bash
# ncu --set full capture of the batched small-system kernel (reduced batch: 16 samples per CTA).
mkdir -p gpurun_out
SHORT="python bench.py --workload c5b --batch 4736 --steps 1 --warmup 1"
timeout -s KILL 300 $SHORT > gpurun_out/plain_c5b.log 2>&1 || { tail -5 gpurun_out/plain_c5b.log; exit 1; }
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k k_batch -s 6 -c 2 -f -o gpurun_out/prof_batch $SHORT > gpurun_out/ncu_batch.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_batch.log
