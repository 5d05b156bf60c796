#!/bin/bash
# Bench lines for the non-default workloads (1 GPU): c5a (shared kappa sweep), c3, c4; plus pipe config alternatives.
mkdir -p gpurun_out
for w in c5a c5b c3 c4; do
  timeout -s KILL 600 python bench.py --workload $w --steps ${STEPS:-3} --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
  python - $w <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
r=d["roofline"]
print(sys.argv[1], "value %.4g %s ms/step %.3f frac %.3f" % (d["value"], d["unit"], d["ms_per_step"], r["frac"]), {k:round(v["ms_per_launch"],4) for k,v in r["kernels"].items()}, r.get("us_per_iteration"))
PY
done
for cfg in ${CFGS:-1 2 3 4}; do
  DFE_PIPE_CFG=$cfg timeout -s KILL 60 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/pm_pipe_$cfg.json 2> gpurun_out/pm_pipe_$cfg.err
  python - "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/pm_pipe_{sys.argv[1]}.json"))
    print("cfg", sys.argv[1], "solves/s %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("cfg", sys.argv[1], "FAILED", e); print(open(f"gpurun_out/pm_pipe_{sys.argv[1]}.err").read()[-800:])
PY
done
