#!/bin/bash
# Pipelined 1-D kernel: wait-primitive experiments (DFE_PIPE_FLAGS) on config 2.
mkdir -p gpurun_out
for fl in ${FLAGS:-0 1 2 3 4 5}; do
  DFE_PIPE_FLAGS=$fl DFE_PIPE_CFG=${CFG:-1} timeout -s KILL 60 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/pf_$fl.json 2> gpurun_out/pf_$fl.err
  python - "$fl" <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/pf_{sys.argv[1]}.json"))
    print("flags", sys.argv[1], "solves/s %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("flags", sys.argv[1], "FAILED", e); print(open(f"gpurun_out/pf_{sys.argv[1]}.err").read()[-800:])
PY
done
