#!/bin/bash
# Pipelined 1-D kernel: parity tests, then config-2 timings per pipeline configuration (short hard timeouts).
mkdir -p gpurun_out
echo "== 1d tests, pipe" > gpurun_out/pipe_check.log
timeout -s KILL ${TEST_TIMEOUT:-240} python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "${TEST_K:-1d}" --timeout 120 --timeout-method=thread -p no:cacheprovider >> gpurun_out/pipe_check.log 2>&1; echo "pipe rc=$?" >> gpurun_out/pipe_check.log
tail -25 gpurun_out/pipe_check.log
for cfg in ${CFGS:-0 1 2 3 4 5 6 7 8}; do
  DFE_1D_MODE=pipe DFE_PIPE_CFG=$cfg timeout -s KILL 60 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/pm_pipe_$cfg.json 2> gpurun_out/pm_pipe_$cfg.err
  python - pipe "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/pm_{sys.argv[1]}_{sys.argv[2]}.json"))
    print(sys.argv[1], sys.argv[2], "solves/s %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e); print(open(f"gpurun_out/pm_{sys.argv[1]}_{sys.argv[2]}.err").read()[-800:])
PY
done
