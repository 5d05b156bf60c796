#!/bin/bash
# stability: the whole -m gpu suite three times, the default bench line twice, the examples
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -1
done
for i in 1 2; do
  timeout -s KILL 900 python bench.py 2>gpurun_out/r2s.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('bench', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'sweep', round(d['sweep']['value']), d['parity']['max_rel'], d['clocks']['reasons'])"
done
timeout -s KILL 600 python examples/poisson_1d_demo.py 2>&1 | tail -4
timeout -s KILL 600 python examples/compliance_2d.py 2>&1 | tail -3
timeout -s KILL 300 python examples/heat_2d.py 2>&1 | tail -3
