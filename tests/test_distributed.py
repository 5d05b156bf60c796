"""Multi-process tests of the batch-sharding layer (gloo, world_size 2, CPU).

The CUDA solver cannot run here, so the local "solver" is the pinned CPU oracle wrapped as a differentiable
torch function: what is tested is the sharding arithmetic and the packed all-reduce — the N > 1 path of
bench.py / config 5 — against the same computation done un-sharded in one process."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from difffe_physics_lab_b200 import distributed as D
from oracle import oracle as O


def test_shard_bounds_tile_the_batch():
    for n in (0, 1, 7, 8, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            b = [D.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(10, 2, 2)
    assert D.world() == (0, 1)
    t = torch.arange(10)
    assert torch.equal(D.shard(t, 1, 3), t[4:7])


class _OracleSolve(torch.autograd.Function):
    """u(kappa) on a line mesh through the CPU oracle (stand-in for the CUDA Function in this CPU test)."""

    @staticmethod
    def forward(ctx, kappa, f, mesh):
        nodes, el, bc = mesh
        u = np.stack([O.forward(nodes, el, bc, float(kappa), fi, exact=False) for fi in f.numpy()])
        ctx.mesh, ctx.kappa = mesh, float(kappa)
        ctx.save_for_backward(torch.from_numpy(u))
        return torch.from_numpy(u)

    @staticmethod
    def backward(ctx, gbar):
        (u,) = ctx.saved_tensors
        nodes, el, bc = ctx.mesh
        g = 0.0
        for ub, gb in zip(u.numpy(), gbar.numpy()):
            gk, _, _ = O.adjoint_and_grads(nodes, el, bc, ctx.kappa, ub, gb, exact=False)
            g += gk.sum()
        return torch.tensor(g, dtype=torch.float64), None, None


def _problem():
    rng = np.random.default_rng(3)
    mesh = O.line_mesh(24)
    n_total = 10
    f = torch.from_numpy(rng.uniform(0.5, 1.5, (n_total, 25)))
    u_data = _OracleSolve.apply(torch.tensor(2.0, dtype=torch.float64), f, mesh)
    return mesh, f, u_data, n_total


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        mesh, f, u_data, n_total = _problem()
        kappa = torch.tensor(1.0, dtype=torch.float64)
        loss, grad = D.sharded_loss_and_grad(lambda k: (lambda fl: _OracleSolve.apply(k, fl, mesh)), kappa,
                                             D.shard(f), D.shard(u_data), n_total)
        # packed all-reduce of several tensors in one collective
        a, b = torch.full((3,), float(rank + 1), dtype=torch.float64), torch.tensor(10.0 * (rank + 1), dtype=torch.float64)
        D.allreduce_sum_([a, b])
        # MisfitSweep: the persistent [dL/dkappa, loss] buffer and its single all-reduce (the local kernels are CUDA-only,
        # so the local step is the oracle here; the product's local_step writes the same two words)
        class OracleSweep(D.MisfitSweep):
            def local_step(self, kap):
                k = kap.detach().clone().requires_grad_(True)
                u = _OracleSolve.apply(k, self.f, self.mesh)
                l = ((u - self.u_data) ** 2).sum() / (u.shape[-1] * self.n_total)
                (g,) = torch.autograd.grad(l, k)
                self.red[0], self.red[1] = g, l.detach()

        sw = OracleSweep(mesh, D.shard(f), D.shard(u_data), n_total)
        buf = sw.red.data_ptr()
        l2, g2 = sw.step(kappa)
        l2, g2 = float(l2), float(g2)
        sw.step(kappa * 1.5)
        assert sw.red.data_ptr() == buf and not sw.registered          # persistent buffer; no NCCL pool on gloo
        out[rank] = (float(loss), float(grad), a.tolist(), float(b), D.shard_bounds(n_total), l2, g2)
    finally:
        dist.destroy_process_group()


def test_sharded_step_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    mesh, f, u_data, n_total = _problem()
    k = torch.tensor(1.0, dtype=torch.float64, requires_grad=True)
    loss = ((_OracleSolve.apply(k, f, mesh) - u_data) ** 2).mean()
    loss.backward()
    for r in (0, 1):
        l, g, a, b, bounds, l2, g2 = out[r]
        assert abs(l - float(loss)) <= 1e-15 + 1e-13 * abs(float(loss))
        assert abs(g - float(k.grad)) <= 1e-13 * abs(float(k.grad))
        assert abs(l2 - float(loss)) <= 1e-15 + 1e-13 * abs(float(loss))
        assert abs(g2 - float(k.grad)) <= 1e-13 * abs(float(k.grad))
        assert a == [3.0, 3.0, 3.0] and b == 30.0
    assert out[0][4] == (0, 5) and out[1][4] == (5, 10)
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]        # identical on every rank
