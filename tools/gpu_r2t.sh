#!/bin/bash
# is the sweep host-bound at the per-rank batch of N = 8?  (8192 samples on one GPU, no collective)
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --workload c5a --batch 8192 --steps 50 --no-cpu 2>gpurun_out/r2t.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c5a B=8192', round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in r['kernels'].items()}, d['gpu_launches'])"
python - <<'P'
import time, torch, sys
sys.path.insert(0, '.')
from difffe_physics_lab_b200 import DifferentiableFESolver, FEMesh
from difffe_physics_lab_b200.distributed import MisfitSweep
dev = torch.device('cuda', 0)
mesh = FEMesh.line(16384)
B = 8192
f = torch.rand((B, 16385), dtype=torch.float64, device=dev) + 0.5
with torch.no_grad():
    u_data = DifferentiableFESolver(mesh, kappa=torch.tensor(2.0, dtype=torch.float64, device=dev))(f)
kappa = torch.tensor(1.0, dtype=torch.float64, device=dev, requires_grad=True)
opt = torch.optim.Adam([kappa], lr=0.05, fused=True)
sw = MisfitSweep(mesh, f, u_data, 65536)
def step():
    loss, grad = sw.step(kappa)
    kappa.grad = grad.detach().reshape(())
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print('host time per step (enqueue only): %.3f ms; wall per step incl. GPU: %.3f ms' % (t_cpu / 200 * 1e3, t_all / 200 * 1e3))
P
