#!/bin/bash
# ncu --set full capture of one k_pcg launch (config 4).
mkdir -p gpurun_out
SHORT="python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout -s KILL 300 $SHORT > gpurun_out/plain_c4.log 2>&1 || { tail -5 gpurun_out/plain_c4.log; exit 1; }
tail -c 400 gpurun_out/plain_c4.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k k_pcg -s 6 -c 1 -f -o gpurun_out/prof_pcg $SHORT > gpurun_out/ncu_pcg.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_pcg.log
