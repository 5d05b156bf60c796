#!/bin/bash
# 2-GPU check of the default bench line (weak c2 + the config-5 sweep with its NCCL all-reduce) and of the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu_n2.txt
export NCCL_DEBUG=WARN
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
tail -5 gpurun_out/bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print('value', d['value'], 'ms', d['ms_per_step'])
print('e2e', d['e2e'] and d['e2e']['value'], 'e2e_full', d.get('e2e_full') and d['e2e_full'].get('value'))
print('sweep', json.dumps(d['sweep'])[:1500])
"
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5a --steps 10 --warmup 3 > gpurun_out/bench_c5a_n2.json 2> gpurun_out/bench_c5a_n2.err; echo "c5a n2 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_c5a_n2.json'))
print('c5a n2', d['value'], d['ms_per_step'], d['config']['collective'])
"
timeout -s KILL 600 python bench.py --workload c5a --steps 10 --warmup 3 > gpurun_out/bench_c5a_n1.json 2> gpurun_out/bench_c5a_n1.err; echo "c5a n1 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_c5a_n1.json'))
print('c5a n1', d['value'], d['ms_per_step'], {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})
"
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"
cat gpurun_out/bench_ref_n2.json | head -c 900
