#!/bin/bash
# 2-D path: parity tests, then config 3 / config 4 bench lines.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not 1d" --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/check2d.log 2>&1; echo "2d tests rc=$?"; tail -3 gpurun_out/check2d.log
for w in c3 c4; do
  timeout -s KILL 600 python bench.py --workload $w --steps ${STEPS:-3} --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
  python - $w <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
r=d["roofline"]
print(sys.argv[1], "value %.4g %s ms/step %.3f frac %.3f" % (d["value"], d["unit"], d["ms_per_step"], r["frac"]), {k:round(v["ms_per_launch"],4) for k,v in r["kernels"].items()}, "us/it", r.get("us_per_iteration"), r.get("iterations"))
PY
done
