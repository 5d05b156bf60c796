#!/bin/bash
# ncu --set full capture of the pipelined 1-D kernels (config 2), after the plain run exited 0.
mkdir -p gpurun_out
SHORT="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
DFE_1D_MODE=pipe DFE_PIPE_CFG=${CFG:-0} timeout -s KILL 120 $SHORT > gpurun_out/plain_pipe.log 2>&1 || { tail -5 gpurun_out/plain_pipe.log; exit 1; }
tail -c 600 gpurun_out/plain_pipe.log
DFE_1D_MODE=pipe DFE_PIPE_CFG=${CFG:-0} timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k k1d_pipe -s 2 -c 2 -f -o gpurun_out/prof_pipe2 $SHORT > gpurun_out/ncu_pipe2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_pipe2.log
