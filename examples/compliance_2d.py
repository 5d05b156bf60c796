#!/usr/bin/env python
"""Compliance gradient with respect to every element's kappa on a 2-D mesh (BASELINE config 4 in miniature; the
reference's README roadmap item "topology-optimisation style objective"): J(kappa) = f^T u, one forward and
one adjoint PCG solve, dJ/dkappa_e for all elements at once, checked against a central finite difference.

    python examples/compliance_2d.py [nx]
"""
import pathlib
import sys
import time

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from diffhe.mesh import FEMesh                      # noqa: E402
from diffhe.solver import DifferentiableFESolver    # noqa: E402


def compliance(mesh, kappa, f):
    """J = sum_i f_i u_i (u = 0 on the boundary): the work of the load, up to the constant nodal weights."""
    u = DifferentiableFESolver(mesh, kappa=kappa)(f)
    return (u * f).sum(), u


def main(nx):
    dev = "cuda"
    mesh = FEMesh.rectangle(nx, nx)
    gen = torch.Generator(device=dev).manual_seed(0)
    kappa = torch.exp(torch.empty(mesh.n_elements, dtype=torch.float64, device=dev).uniform_(-3.0, 0.0, generator=gen))
    kappa.requires_grad_(True)
    f = torch.ones(mesh.n_nodes, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    J, _ = compliance(mesh, kappa, f)
    J.backward()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    g = kappa.grad.clone()
    print(f"rectangle({nx},{nx}): {mesh.n_elements} elements, J = {float(J.detach()):.12e}, |dJ/dkappa|_1 = {float(g.abs().sum()):.6e}, "
          f"forward + adjoint in {dt * 1e3:.1f} ms")
    # finite-difference check of a few components
    with torch.no_grad():
        for e in (mesh.n_elements // 3, mesh.n_elements // 2 + nx, (2 * mesh.n_elements) // 3 + 7):
            h = 1e-3 * float(kappa[e])
            kp, km = kappa.detach().clone(), kappa.detach().clone()
            kp[e] += h
            km[e] -= h
            fd = (float(compliance(mesh, kp, f)[0]) - float(compliance(mesh, km, f)[0])) / (2 * h)
            print(f"  element {e}: adjoint {float(g[e]):+.9e}   central difference {fd:+.9e}")


if __name__ == "__main__":
    if not torch.cuda.is_available():
        raise SystemExit("this example needs a CUDA device (difffe_physics_lab_b200 has no CPU fallback)")
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 256)
