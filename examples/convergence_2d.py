"""Convergence of the accelerated path on -laplace(u) = 2 pi^2 sin(pi x) sin(pi y), u = 0 on the boundary of the unit square:
the reference's P1 elements against quadratic (P2) elements on the same vertices.

Two roadmap items of the reference (README.md:139-143: "2D convergence tests", "P2 elements") on top of the same module:
``FEMesh.rectangle(n, n)`` is the reference's mesh, ``.to_p2()`` adds the edge midpoints; ``DifferentiableFESolver`` then
assembles the P2 stiffness matrix and the consistent load (csrc/dfe_p2.cuh) and solves with the same CSR / PCG kernels.
Printed: nodal max error and observed order (P1 with the reference's centroid load rule: 2; P2: >= 3), and a gradient
check of d(sum u)/dkappa = -sum(u)/kappa for P2.

    python examples/convergence_2d.py            (needs a CUDA device)
"""
import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from difffe_physics_lab_b200 import DifferentiableFESolver, FEMesh   # noqa: E402


def error(mesh):
    x, y = mesh.nodes[:, 0], mesh.nodes[:, 1]
    exact = torch.sin(np.pi * x) * torch.sin(np.pi * y)
    u = DifferentiableFESolver(mesh)((2 * np.pi ** 2 * exact).cuda()).cpu()
    return float((u - exact).abs().max())


def main():
    prev = None
    print(f"{'n':>5} {'P1 error':>12} {'order':>6} {'P2 error':>12} {'order':>6} {'P2 unknowns':>12}")
    for n in (8, 16, 32, 64, 128):
        m1 = FEMesh.rectangle(n, n)
        m2 = m1.to_p2()
        e1, e2 = error(m1), error(m2)
        o = ("", "") if prev is None else (f"{np.log2(prev[0] / e1):.2f}", f"{np.log2(prev[1] / e2):.2f}")
        print(f"{n:>5} {e1:>12.3e} {o[0]:>6} {e2:>12.3e} {o[1]:>6} {m2.n_nodes - len(m2.dirichlet_nodes):>12}")
        prev = (e1, e2)
    assert np.log2(prev[0] / e1) >= 0 and e2 < 1e-3 * e1
    m = FEMesh.rectangle(24, 24).to_p2()
    kappa = torch.tensor(1.7, dtype=torch.float64, device="cuda", requires_grad=True)
    u = DifferentiableFESolver(m, kappa=kappa)(torch.ones(m.n_nodes, dtype=torch.float64, device="cuda"))
    u.sum().backward()
    rel = abs(float(kappa.grad) + float(u.sum()) / 1.7) / abs(float(u.sum()) / 1.7)
    print(f"P2 gradient check: d(sum u)/dkappa = {float(kappa.grad):.12e}, -sum(u)/kappa = {-float(u.sum()) / 1.7:.12e} (rel {rel:.1e})")
    assert rel < 1e-9


if __name__ == "__main__":
    main()
