#!/bin/bash
# Run on the GPU box via gpurun: default bench line, then (only if it exited 0) the ncu launch list and
# one --set full capture of the 1-D kernels.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
rc=$?; echo "bench rc=$rc"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
[ $rc -ne 0 ] && exit $rc
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout -s KILL 600 $SHORT > gpurun_out/plain.log 2>&1 &&
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
echo "launches rc=$?"
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:k_solve1d -s 6 -c 2 -f -o gpurun_out/prof_1d $SHORT > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -5 gpurun_out/ncu_full.log
