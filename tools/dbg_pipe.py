"""Debug helper (GPU box): pipelined 1-D kernel vs the CPU oracle on a few shapes, per-sample error table."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffhe.mesh import FEMesh
from diffhe.solver import DifferentiableFESolver
from oracle import oracle as O

def run(mesh, kappa, f, gbar):
    k = torch.as_tensor(kappa, device="cuda").requires_grad_(True)
    ft = torch.as_tensor(f, device="cuda").requires_grad_(True)
    u = DifferentiableFESolver(mesh, kappa=k)(ft)
    (u * torch.as_tensor(gbar, device="cuda")).sum().backward()
    return u.detach().cpu().numpy(), k.grad.cpu().numpy(), ft.grad.cpu().numpy()

cases = [(5000, 4, (0.5, None)), (5000, 1, (0.5, None)), (5000, 4, (0.0, None)), (1500, 4, (0.5, None)),
         (5000, 4, (None, 0.5)), (20000, 3, (None, 1.5)), (100000, 3, (0.0, 0.0)), (100000, 40, (0.25, -0.5))]
for n, B, bcs in cases:
    rng = np.random.default_rng(n)
    m = FEMesh.line(n, x_left=-0.2, x_right=1.3, bc_left=bcs[0], bc_right=bcs[1])
    f = rng.uniform(0, 1, (B, n + 1)); gbar = rng.standard_normal((B, n + 1))
    kap = np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1)))
    u, gk, gf = run(m, kap, f, gbar)
    nodes, el, bc = m.nodes.numpy(), m.elements.numpy(), m.dirichlet_nodes
    for b in range(min(B, 4)):
        uo = O.forward(nodes, el, bc, float(kap[b, 0]), f[b])
        gko, gfo, _ = O.adjoint_and_grads(nodes, el, bc, float(kap[b, 0]), uo, gbar[b])
        eu = np.abs(u[b] - uo).max() / np.abs(uo).max()
        egk = abs(gk[b, 0] - gko.sum()) / np.abs(gko).sum()
        egf = np.abs(gf[b] - gfo).max() / np.abs(gfo).max()
        print(f"n={n} B={B} bcs={bcs} b={b}: eu={eu:.2e} egk={egk:.2e} egf={egf:.2e}  gk={gk[b,0]:.6f} ref={gko.sum():.6f} "
              f"diff*kap={(gk[b,0]-gko.sum())*kap[b,0]:.6f} sum(gbar)={gbar[b].sum():.6f} sum(gbar[1:])={gbar[b,1:].sum():.6f}", flush=True)
