#!/bin/bash
# ncu evidence for profiles/: launch lists of the default bench (config 2) and of config 4, plus --set full
# captures of the dominant kernels.  Each ncu pass only after the same command exited 0 without ncu.
mkdir -p gpurun_out
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout -s KILL 600 $SHORT > gpurun_out/plain_c2.log 2>&1 &&
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv $SHORT > /dev/null 2>&1
echo "launches c2 rc=$?"
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:k1d_pass -s 12 -c 4 -f -o gpurun_out/prof_c2 $SHORT > gpurun_out/ncu_c2.log 2>&1
echo "full c2 rc=$?"
S4="python bench.py --workload c4 --steps 1 --warmup 3"
timeout -s KILL 600 $S4 > gpurun_out/plain_c4.log 2>&1 &&
timeout -s KILL 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4.csv $S4 > /dev/null 2>&1
echo "launches c4 rc=$?"
