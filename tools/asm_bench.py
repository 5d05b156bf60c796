"""Time dfe_assemble on rectangle(n, n) with per-element kappa (config 4 geometry): CUDA events over back-to-back launches
(warm L2) and with an L2 flush between launches.  Usage: python tools/asm_bench.py [n] [reps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from difffe_physics_lab_b200 import FEMesh, _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda", 0)
m = FEMesh.rectangle(n, n)
nm = m._native(0)
I = nm.info
L = _native.lib()
g = torch.Generator(device="cpu").manual_seed(0)
kap = torch.exp(torch.rand(m.n_elements, dtype=torch.float64, generator=g) * np.log(1e-3)).to(dev)
f = torch.ones(m.n_nodes, dtype=torch.float64, device=dev)
vals = torch.empty(I.nnz_full, dtype=torch.float64, device=dev)
F = torch.empty(I.n_nodes, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def launch():
    _native.check(L.dfe_assemble(nm.handle, kap.data_ptr(), _native.KAPPA_PER_ELEMENT, f.data_ptr(), vals.data_ptr(), F.data_ptr(), st))

for _ in range(5):
    launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    launch()
e1.record()
torch.cuda.synchronize()
warm = e0.elapsed_time(e1) / reps * 1e3
cold = []
for _ in range(reps):
    flush.fill_(1)
    e0.record()
    launch()
    e1.record()
    torch.cuda.synchronize()
    cold.append(e0.elapsed_time(e1) * 1e3)
alg = 12 * I.n_elements + 16 * I.n_nodes + 8 * I.n_elements + 8 * I.nnz_free + 16 * I.n_nodes
print(json.dumps({"n": n, "env": {k: v for k, v in os.environ.items() if k.startswith("DFE_ASSEMBLE")}, "warm_us": round(warm, 2),
                  "cold_us_median": round(float(np.median(cold)), 2), "cold_us_min": round(min(cold), 2), "alg_MB": round(alg / 1e6, 1),
                  "frac_cold": round(alg / (np.median(cold) * 1e-6) / 6533.2e9, 3), "checksum": float(vals.sum().item())}))
