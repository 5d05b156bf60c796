#!/bin/bash
# Compare the 1-D kernel variants / slab sizes on the config-2 workload (device-resident timing only).
mkdir -p gpurun_out
for cfg in "split 0" "split 32" "split 64" "split 96" "seq 0"; do
  set -- $cfg
  DFE_1D_MODE=$1 DFE_1D_SLAB_MB=$2 timeout -s KILL 300 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/mode_$1_$2.json 2> gpurun_out/mode_$1_$2.err
  python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/mode_{sys.argv[1]}_{sys.argv[2]}.json"))
    print(sys.argv[1], sys.argv[2], "solves/s %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e); print(open(f"gpurun_out/mode_{sys.argv[1]}_{sys.argv[2]}.err").read()[-800:])
PY
done
