#!/bin/bash
# Timeline of the pipelined 1-D kernel (clock64 stamps per role, CTAs 0/100/200, iterations 200..215).
mkdir -p gpurun_out
DFE_PIPE_TRACE=1 DFE_PIPE_CFG=${CFG:-0} timeout -s KILL 120 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/trace_run.log 2>&1
tail -c 400 gpurun_out/trace_run.log
