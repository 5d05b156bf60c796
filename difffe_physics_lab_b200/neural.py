"""``NeuralPDE`` — small MLP PDE surrogate, kept API-compatible (mirror of reference ``diffhe/neural.py``).

Out of the accelerated scope: a float64 tanh MLP over the mesh nodes whose output is
multiplied by a mask vanishing on the Dirichlet nodes (``diffhe/neural.py:19-101``), trained with Adam
against ``PhysicsLoss`` (``neural.py:105-149``).  It reaches the CUDA hot path only through
``PhysicsLoss(mode="fem_match")``.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.nn as nn

from .loss import PhysicsLoss
from .mesh import FEMesh


class NeuralPDE(nn.Module):
    """``u(x) = mask(x) * MLP(x)`` with ``n_layers`` tanh hidden layers of width ``hidden_dim``."""

    def __init__(self, mesh: FEMesh, hidden_dim: int = 32, n_layers: int = 3):
        super().__init__()
        self.mesh = mesh
        self.dim = mesh.dim
        widths = [self.dim] + [hidden_dim] * n_layers
        blocks: List[nn.Module] = []
        for w_in, w_out in zip(widths[:-1], widths[1:]):
            blocks.append(nn.Linear(w_in, w_out))
            blocks.append(nn.Tanh())
        blocks.append(nn.Linear(hidden_dim, 1))
        self.net = nn.Sequential(*blocks).double()
        self._mask = self._compute_mask()

    def forward(self, x: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Network solution at ``x`` (default: the mesh nodes), shape ``(n_nodes,)``."""
        pts = (self.mesh.nodes if x is None else x).double()
        out = self.net(pts).squeeze(1)
        return self._mask.to(pts.device) * out

    def _compute_mask(self) -> torch.Tensor:
        """0 on Dirichlet nodes; 1D with two constrained ends: the parabola (x-a)(b-x) scaled to max ~1."""
        nodes = self.mesh.nodes
        n = nodes.shape[0]
        fixed = list(self.mesh.dirichlet_nodes.keys())
        if self.dim == 1:
            if len(fixed) < 2:
                return torch.ones(n, dtype=torch.float64)
            x = nodes[:, 0]
            a, b = float(nodes[fixed[0], 0]), float(nodes[fixed[-1], 0])
            bump = (x - a) * (b - x)
            return bump / (bump.abs().max() + 1e-12)
        mask = torch.ones(n, dtype=torch.float64)
        if fixed:
            mask[torch.tensor(fixed, dtype=torch.long)] = 0.0
        return mask

    def train_pde(self, forcing_fn: Callable[[torch.Tensor], torch.Tensor], n_epochs: int = 2000,
                  lr: float = 1e-3, mode: str = "fem_match", verbose: bool = True,
                  log_every: int = 200) -> List[float]:
        """Adam on ``PhysicsLoss(mesh, forcing_fn, mode)``; returns the loss history."""
        criterion = PhysicsLoss(self.mesh, forcing_fn, mode=mode)
        opt = torch.optim.Adam(self.parameters(), lr=lr)
        history: List[float] = []
        for epoch in range(1, n_epochs + 1):
            opt.zero_grad()
            loss = criterion(self.forward())
            loss.backward()
            opt.step()
            history.append(float(loss))
            if verbose and epoch % log_every == 0:
                print(f"  Epoch {epoch:5d}  loss = {history[-1]:.3e}")
        return history
