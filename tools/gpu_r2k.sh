#!/bin/bash
mkdir -p gpurun_out
SHORT5="python bench.py --workload c5b --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_band_rhs_fwd3|k_band_grad3" -s 2 -c 2 -f -o gpurun_out/prof_band3 $SHORT5 > gpurun_out/ncu_band3.log 2>&1; echo "ncu band3 rc=$?"
