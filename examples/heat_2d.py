"""Heat equation du/dt = div(kappa grad u) on the unit square, backward Euler, for a batch of initial states at once.

A roadmap item of the reference (README.md:139-143) built on the accelerated path: ``M_L + dt K`` is assembled with the
bit-exact P1 kernels, factored ONCE by the banded Cholesky and reused by every step (block TRSM on the FP64 tensor cores);
see difffe_physics_lab_b200/timestepping.py.

Checks printed: (1) the first eigenmode decays by exactly 1 / (1 + dt lambda_h) per step, lambda_h the eigenvalue of the
lumped-mass 5-point operator (the discrete scheme has a closed form on the uniform mesh); (2) random initial states against
a sparse direct solve of the same step on the CPU (scipy); (3) throughput for 4096 states.

    python examples/heat_2d.py            (needs a CUDA device)
"""
import pathlib
import sys
import time

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from difffe_physics_lab_b200 import FEMesh                      # noqa: E402
from difffe_physics_lab_b200.timestepping import HeatStepper     # noqa: E402


def main():
    nx, kappa, dt, nsteps = 32, 0.7, 2e-3, 50
    mesh = FEMesh.rectangle(nx, nx)
    hs = HeatStepper(mesh, kappa, dt)
    x, y = mesh.nodes[:, 0].cuda(), mesh.nodes[:, 1].cuda()
    # (1) first eigenmode
    u0 = (torch.sin(np.pi * x) * torch.sin(np.pi * y)).reshape(1, -1)
    u = hs.step(u0, nsteps)
    h = 1.0 / nx
    lam = kappa * 8.0 / h ** 2 * np.sin(np.pi * h / 2) ** 2
    decay = (1.0 + dt * lam) ** (-nsteps)
    err = float((u - decay * u0).abs().max() / (decay * u0.abs().max()))
    print(f"eigenmode after {nsteps} steps: decay {float(u.max() / u0.max()):.12f}, discrete closed form {decay:.12f} "
          f"(rel. error {err:.2e}), continuum exp(-2 pi^2 kappa t) = {np.exp(-2 * np.pi ** 2 * kappa * dt * nsteps):.6f}")
    assert err < 1e-10
    # (2) random states, heterogeneous kappa, non-zero boundary value, source term: against scipy
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    rng = np.random.default_rng(0)
    mesh2 = FEMesh.rectangle(24, 17, x_range=(0.0, 1.4), y_range=(-0.3, 0.5), bc_value=0.2)
    kap_e = np.exp(rng.uniform(np.log(0.05), 0.0, mesh2.n_elements))
    fsrc = torch.tensor(rng.uniform(-1, 1, mesh2.n_nodes))
    hs2 = HeatStepper(mesh2, kap_e, 5e-3, f=fsrc)
    B = 6
    u0 = torch.tensor(rng.uniform(-1, 1, (B, mesh2.n_nodes)), device="cuda")
    bc_idx = torch.tensor(list(mesh2.dirichlet_nodes), device="cuda")
    u0[:, bc_idx] = 0.2
    u = hs2.step(u0, 3).cpu().numpy()
    rp, col = hs2.nm.csr(0)
    A = sp.csr_matrix((hs2.A.cpu().numpy(), col, rp), shape=(mesh2.n_nodes,) * 2)
    free = hs2.free.cpu().numpy()
    lu = spla.splu(A[free][:, free].tocsc())
    ref = u0.cpu().numpy()
    mass, load = hs2.mass.cpu().numpy(), hs2.load.cpu().numpy()
    for _ in range(3):
        rhs = mass * ref + 5e-3 * load - (A @ hs2.g.cpu().numpy())
        nxt = np.tile(hs2.g.cpu().numpy(), (B, 1))
        nxt[:, free] = lu.solve(rhs[:, free].T).T
        ref = nxt
    err = np.abs(u - ref).max() / np.abs(ref).max()
    print(f"3 steps, per-element kappa, source, bc 0.2, {B} states vs scipy splu: max rel. diff {err:.2e}")
    assert err < 1e-10
    # (3) throughput
    B = 4096
    u0 = torch.rand((B, mesh.n_nodes), dtype=torch.float64, device="cuda")
    hs.step(u0, 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hs.step(u0, 100)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(f"{B} states x 100 implicit steps on rectangle({nx},{nx}) (961 unknowns): {t * 1e3:.1f} ms = "
          f"{B * 100 / t / 1e6:.2f} M state-steps/s, one factorisation")


if __name__ == "__main__":
    main()
