"""Backward-Euler time stepping of the heat equation with ONE banded factorisation reused by every step.

A caller of the accelerated path from the reference's roadmap (README.md:139-143: "time-dependent heat equation"; SURVEY §8f
N4) — nothing of this exists upstream, so there is no parity burden beyond the operators it is built from:

    M_L du/dt + K(kappa) u = F(f),        (M_L + dt K) u^{n+1} = M_L u^n + dt F

with K and F exactly what ``DifferentiableFESolver`` assembles (``dfe_assemble``: reference solver.py:112-145, bit for bit) and
M_L the lumped mass matrix — the load vector of f = 1 (solver.py:143-145).  ``dfe_band_factor`` factors ``M_L + dt K`` once
(any SPD matrix on the pattern of K is accepted), ``dfe_band_solve`` applies the two triangular solves to a whole batch of
states per step (block TRSM on the FP64 tensor cores).  Meshes: half bandwidth of K_free <= 32 (``rectangle(nx, ny)``, nx <= 32).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native
from .mesh import FEMesh


class HeatStepper:
    """``step(u)`` advances a batch of states ``u`` (B, n_nodes) by ``dt``; Dirichlet nodes keep their boundary values."""

    def __init__(self, mesh: FEMesh, kappa, dt: float, f: torch.Tensor | None = None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("HeatStepper runs on CUDA only (difffe_physics_lab_b200 has no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.mesh, self.dt, self.dev = mesh, float(dt), dev
        L = _native.lib()
        self.nm = nm = mesh._native(dev.index)
        if not L.dfe_band_supported(nm.handle):
            raise NotImplementedError("HeatStepper needs a mesh whose K_free has half bandwidth <= 32 (dfe_band_supported)")
        I = nm.info
        kap = torch.as_tensor(kappa, dtype=torch.float64, device=dev).reshape(-1).contiguous()
        mode = _native.KAPPA_SCALAR if kap.numel() == 1 else _native.KAPPA_PER_ELEMENT
        st = torch.cuda.current_stream(dev).cuda_stream
        ones = torch.ones(I.n_nodes, dtype=torch.float64, device=dev)
        vals = torch.empty(I.nnz_full, dtype=torch.float64, device=dev)
        mass = torch.empty(I.n_nodes, dtype=torch.float64, device=dev)
        _native.check(L.dfe_assemble(nm.handle, kap.data_ptr(), mode, ones.data_ptr(), vals.data_ptr(), mass.data_ptr(), st))
        self.load = torch.zeros(I.n_nodes, dtype=torch.float64, device=dev)
        if f is not None:                                   # F(f): the reference's load vector of the source term
            fs = f.to(device=dev, dtype=torch.float64).contiguous()
            scratch = torch.empty_like(vals)
            _native.check(L.dfe_assemble(nm.handle, kap.data_ptr(), mode, fs.data_ptr(), scratch.data_ptr(), self.load.data_ptr(), st))
        rp, col = nm.csr(0)
        rows = np.repeat(np.arange(I.n_nodes), np.diff(rp))
        diag = torch.as_tensor(np.nonzero(col == rows)[0], device=dev)
        self.A = self.dt * vals                             # M_L + dt K on the pattern of K
        self.A[diag] += mass
        self.K_vals, self.mass = vals, mass
        self.factor = torch.empty(L.dfe_band_factor_bytes(nm.handle), dtype=torch.uint8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _native.check(L.dfe_band_factor(nm.handle, self.A.data_ptr(), self.factor.data_ptr(), status.data_ptr(), st))
        if int(status) != 0:
            raise _native.BreakdownError(_native.ERR_BREAKDOWN, "M_L + dt K is not positive definite")
        self.free = torch.as_tensor(np.asarray(mesh.free_nodes(), dtype=np.int64), device=dev)
        self.npad = int(L.dfe_band_npad(nm.handle))
        # Dirichlet columns of (M_L + dt K) times the boundary values: the lifting of the implicit step, per free row
        bc = mesh.dirichlet_nodes
        g = torch.zeros(I.n_nodes, dtype=torch.float64, device=dev)
        if bc:
            g[torch.as_tensor(list(bc.keys()), device=dev)] = torch.as_tensor(list(bc.values()), dtype=torch.float64, device=dev)
        self.g = g
        crow = torch.as_tensor(rows, device=dev)
        ccol = torch.as_tensor(col, device=dev)
        lift = torch.zeros(I.n_nodes, dtype=torch.float64, device=dev)
        lift.index_add_(0, crow, self.A * g[ccol])          # (A g)_p; only Dirichlet columns contribute (g = 0 elsewhere)
        self.rhs_const = (self.dt * self.load - lift)[self.free]
        self.mass_free = mass[self.free]
        self._X = None

    def step(self, u: torch.Tensor, n_steps: int = 1) -> torch.Tensor:
        """u (B, n_nodes) float64 on the stepper's device -> the states after ``n_steps`` steps (new tensor)."""
        L = _native.lib()
        B = u.shape[0]
        nf = self.free.numel()
        if self._X is None or self._X.shape[0] != B:
            self._X = torch.zeros((B, self.npad), dtype=torch.float64, device=self.dev)
        X = self._X
        X[:, :nf] = u[:, self.free]
        st = torch.cuda.current_stream(self.dev).cuda_stream
        for _ in range(n_steps):
            X[:, :nf].mul_(self.mass_free).add_(self.rhs_const)                 # M_L u^n + dt F - lifting
            _native.check(L.dfe_band_solve(self.nm.handle, B, X.data_ptr(), self.factor.data_ptr(), st))
        out = self.g.expand(B, -1).clone()
        out[:, self.free] = X[:, :nf]
        return out
