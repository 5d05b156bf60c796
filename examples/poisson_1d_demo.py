#!/usr/bin/env python
"""1-D Poisson demo on the B200 path — the three parts of the reference's examples/poisson_1d_demo.py, written against
the same public API (`from diffhe... import ...` resolves to this repository's package when it is first on sys.path):

  1. FEM solve of -u'' = 1 on line(20) against the exact solution x(1-x)/2,
  2. autograd gradient d(sum u)/dkappa against the analytic value -sum(u)/kappa,
  3. recovery of kappa = 2 from data by 200 Adam steps through the solver,
plus what the reference cannot do: the same recovery for 4096 samples with their own kappa in one batched call.

Needs a CUDA device (there is no CPU fallback):  python examples/poisson_1d_demo.py
"""
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from diffhe.mesh import FEMesh                      # noqa: E402
from diffhe.solver import DifferentiableFESolver    # noqa: E402


def main():
    dev = "cuda"
    # ---- 1. forward solve
    mesh = FEMesh.line(n_elements=20)
    f = torch.ones(mesh.n_nodes, dtype=torch.float64, device=dev)
    u = DifferentiableFESolver(mesh, kappa=1.0)(f)
    x = mesh.nodes[:, 0].to(dev)
    print(f"1. line(20), kappa = 1, f = 1: max |u - x(1-x)/2| = {float((u - 0.5 * x * (1 - x)).abs().max()):.2e}")

    # ---- 2. gradient with respect to kappa
    kappa = torch.tensor(1.0, dtype=torch.float64, device=dev, requires_grad=True)
    total = DifferentiableFESolver(mesh, kappa=kappa)(f).sum()
    total.backward()
    print(f"2. d(sum u)/dkappa = {float(kappa.grad):.16f}   (analytic -sum(u)/kappa = {-float(total.detach()):.16f})")

    # ---- 3. recover kappa from data
    mesh = FEMesh.line(n_elements=30)
    f = torch.ones(mesh.n_nodes, dtype=torch.float64, device=dev)
    with torch.no_grad():
        u_data = DifferentiableFESolver(mesh, kappa=2.0)(f)
    k = torch.tensor(1.0, dtype=torch.float64, device=dev, requires_grad=True)
    opt = torch.optim.Adam([k], lr=0.1)
    for _ in range(200):
        opt.zero_grad()
        loss = ((DifferentiableFESolver(mesh, kappa=k.abs())(f) - u_data) ** 2).mean()
        loss.backward()
        opt.step()
    print(f"3. recovered kappa = {float(k.detach().abs()):.4f}   (true 2.0000), final loss {float(loss.detach()):.2e}")

    # ---- 4. the batched form: 4096 independent recoveries in one call per step
    B = 4096
    mesh = FEMesh.line(n_elements=2000)
    gen = torch.Generator(device=dev).manual_seed(0)
    fb = torch.rand((B, mesh.n_nodes), dtype=torch.float64, device=dev, generator=gen) + 0.5
    k_true = torch.exp(torch.empty((B, 1), dtype=torch.float64, device=dev).uniform_(-0.7, 0.7, generator=gen))
    with torch.no_grad():
        ub = DifferentiableFESolver(mesh, kappa=k_true)(fb)
    logk = torch.zeros((B, 1), dtype=torch.float64, device=dev, requires_grad=True)
    opt = torch.optim.Adam([logk], lr=0.05)
    for _ in range(300):
        opt.zero_grad()
        # per-sample losses, scaled so that each sample sees an O(1) problem
        loss = (((DifferentiableFESolver(mesh, kappa=logk.exp())(fb) - ub) / ub.abs().amax(dim=1, keepdim=True)) ** 2).mean(dim=1).sum()
        loss.backward()
        opt.step()
    err = float(((logk.exp() - k_true).abs() / k_true).max())
    print(f"4. {B} samples on line(2000), per-sample kappa: max relative error of the recovered kappa = {err:.2e}")


if __name__ == "__main__":
    if not torch.cuda.is_available():
        raise SystemExit("this demo needs a CUDA device (difffe_physics_lab_b200 has no CPU fallback)")
    main()
