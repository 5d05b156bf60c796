#!/usr/bin/env python
"""Topology optimisation on the accelerated path: distribute a limited amount of conductive material so that a uniformly
heated plate, cooled along its whole boundary, stays as cold as possible (thermal compliance, SIMP + optimality criteria).

The reference's roadmap item "Topology optimisation demo (minimise compliance)" (README.md:139-143 upstream) written as a
caller of the unchanged module API — the workload of BASELINE config 4 in a loop:

    kappa_e = k_min + (1 - k_min) rho_e^p            element conductivities from the design densities rho_e in [0, 1]
    C(rho)  = sum_i f_i u_i                          thermal compliance, u = DifferentiableFESolver(mesh, kappa)(f)
    dC/drho                                          one adjoint solve (autograd through the solver)
    rho <- OC update under  mean(rho) = volume fraction,  sensitivities smoothed by a 3 x 3 quad filter

Every iteration is one forward and one adjoint solve with per-element kappa of contrast 1e3 on `rectangle(n, n)`: assembly
(bit-exact tile kernel), multigrid-preconditioned CG, element-gradient kernel.  Printed: compliance, volume and solver
iterations per design step; the script asserts that the compliance falls and the volume constraint holds.

    python examples/topopt_heat.py [n] [steps]            (needs a CUDA device)
"""
import pathlib
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from diffhe.mesh import FEMesh                      # noqa: E402
from diffhe.solver import DifferentiableFESolver    # noqa: E402


def smooth(g, n):
    """3 x 3 box filter on the quad grid (both triangles of a quad share the value): mesh-independent sensitivities."""
    q = g.reshape(n, n, 2).mean(dim=2)[None, None]
    q = F.avg_pool2d(F.pad(q, (1, 1, 1, 1), mode="replicate"), 3, stride=1)[0, 0]
    return q[:, :, None].expand(n, n, 2).reshape(-1)


def oc_update(rho, dc, vol, move=0.15):
    """Optimality-criteria update: rho * sqrt(-dC/drho / lambda) clipped to the move limit, lambda by bisection on the volume."""
    lo, hi = 1e-12, 1e12
    for _ in range(80):
        lam = (lo * hi) ** 0.5
        new = (rho * torch.sqrt(torch.clamp(-dc, min=0.0) / lam)).clamp(min=1e-3, max=1.0)
        new = torch.minimum(torch.maximum(new, rho - move), rho + move).clamp(1e-3, 1.0)
        if float(new.mean()) > vol:
            lo = lam
        else:
            hi = lam
    return new


def main(n, steps):
    dev = "cuda"
    mesh = FEMesh.rectangle(n, n)
    f = torch.ones(mesh.n_nodes, dtype=torch.float64, device=dev)
    vol, p, kmin = 0.3, 3.0, 1e-3
    rho = torch.full((mesh.n_elements,), vol, dtype=torch.float64, device=dev)
    hist = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(steps):
        rho.requires_grad_(True)
        kappa = kmin + (1.0 - kmin) * rho ** p
        solver = DifferentiableFESolver(mesh, kappa=kappa)
        u = solver(f)
        c = (u * f).sum()
        c.backward()
        dc = smooth(rho.grad, n)
        its = solver.last_pcg[0][0]
        hist.append(float(c.detach()))
        if it % 5 == 0 or it == steps - 1:
            print(f"step {it:3d}: compliance {hist[-1]:.6e}  volume {float(rho.detach().mean()):.4f}  "
                  f"solid fraction (rho > 0.9) {float((rho.detach() > 0.9).double().mean()):.3f}  CG iterations {its}")
        rho = oc_update(rho.detach(), dc, vol)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{steps} design steps on rectangle({n},{n}) ({mesh.n_elements} design variables): {dt:.2f} s, "
          f"{1e3 * dt / steps:.1f} ms per step (forward + adjoint + update); compliance {hist[0]:.4e} -> {hist[-1]:.4e}")
    assert hist[-1] < 0.5 * hist[0], "the optimised design should at least halve the compliance of the uniform one"
    assert abs(float(rho.mean()) - vol) < 2e-3


if __name__ == "__main__":
    if not torch.cuda.is_available():
        raise SystemExit("this example needs a CUDA device (difffe_physics_lab_b200 has no CPU fallback)")
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 40)
