#!/bin/bash
# N-GPU check of the default bench line (weak c2 + the config-5 sweep with its NCCL all-reduce), c5b, and the reference arm
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu_n$N.txt
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
lscpu | grep -E "^CPU\(s\)|NUMA|Model name" > gpurun_out/lscpu_n$N.txt
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"
python - <<P
import json
def load(fn):
    for ln in open(fn):
        if ln.startswith('{'): return json.loads(ln)
d=load('gpurun_out/bench_n$N.json')
print('value', d['value'], 'ms', d['ms_per_step'])
print('e2e', d['e2e'] and d['e2e']['value'], 'e2e_full', d.get('e2e_full') and d['e2e_full'].get('value'))
s=d['sweep']; print('sweep', s['value'], s['ms_per_step'], s['collective'], {k:round(v['ms_per_launch'],3) for k,v in s['kernels'].items()})
P
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c5b --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_c5b_n$N.json 2> gpurun_out/bench_c5b_n$N.err; echo "c5b n$N rc=$?"
python - <<P
import json
def load(fn):
    for ln in open(fn):
        if ln.startswith('{'): return json.loads(ln)
d=load('gpurun_out/bench_c5b_n$N.json')
print('c5b n$N', d['value'], d['ms_per_step'], {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})
P
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref n$N rc=$?"
head -c 400 gpurun_out/bench_ref_n$N.json; echo
