"""GPU tests of the P2 extension (reference roadmap item; no upstream arithmetic exists): the CUDA path through the module
API and the C ABI against the quadrature-based oracle (oracle/oracle_p2.py, validated on the CPU against closed-form
solutions) and against closed-form solutions directly.  Tolerance: 1e-9 relative after PCG to a 1e-13 residual, as for
the P1 2-D path; assembled values 1e-13."""
import numpy as np
import pytest
import torch

from difffe_physics_lab_b200 import _native
from diffhe.mesh import FEMesh
from diffhe.solver import DifferentiableFESolver
from oracle import oracle_p2 as P

from .test_gpu_parity import abi_assemble, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _native.build()


def meshes():
    rng = np.random.default_rng(7)
    m1 = FEMesh.line(16, 0.2, 1.7, bc_left=0.4, bc_right=-0.3).to_p2()
    m1n = FEMesh.line(9, bc_left=1.0, bc_right=None).to_p2()                    # natural right end
    m2 = FEMesh.rectangle(9, 7, x_range=(-0.5, 1.1), y_range=(0.0, 0.8), bc_value=0.25).to_p2()
    m2p = FEMesh.rectangle(6, 5).to_p2()
    for k in list(m2p.dirichlet_nodes)[::3]:                                     # partial Dirichlet set
        del m2p.dirichlet_nodes[k]
    m2j = FEMesh.rectangle(5, 6)                                                 # jittered interior vertices
    x = m2j.nodes.clone()
    interior = [p for p in range(m2j.n_nodes) if p not in m2j.dirichlet_nodes]
    x[interior] += torch.from_numpy(rng.uniform(-0.04, 0.04, (len(interior), 2)))
    m2j = FEMesh(nodes=x, elements=m2j.elements, dirichlet_nodes=m2j.dirichlet_nodes).to_p2()
    return {"line16": m1, "line9_natural": m1n, "rect9x7": m2, "rect6x5_partial": m2p, "rect5x6_jitter": m2j}


@pytest.mark.parametrize("name", ["line16", "line9_natural", "rect9x7", "rect6x5_partial", "rect5x6_jitter"])
@pytest.mark.parametrize("per_elem", [False, True])
def test_p2_vs_oracle(name, per_elem):
    m = meshes()[name]
    rng = np.random.default_rng(len(name) + per_elem)
    nodes, el, bc = m.nodes.numpy(), m.elements.numpy(), m.dirichlet_nodes
    kap = np.exp(rng.uniform(-1.5, 1.0, m.n_elements)) if per_elem else np.asarray(1.3)
    f = rng.uniform(-1, 1, m.n_nodes)
    g = rng.normal(size=m.n_nodes)
    # assembled K, F through the ABI
    nm, vals, F, *_ = abi_assemble(m, kap, f)
    rp, col = nm.csr(0)
    Ko, Fo, _, _ = P.assemble(nodes, el, kap, f)
    Kd = np.asarray(Ko.todense())
    rows = np.repeat(np.arange(m.n_nodes), np.diff(rp))
    assert np.abs(vals - Kd[rows, col]).max() <= 1e-13 * np.abs(Kd).max()
    assert np.abs(F - Fo).max() <= 1e-13 * np.abs(Fo).max()
    # u and gradients through the module
    k = torch.as_tensor(kap, device="cuda").requires_grad_(True)
    ft = torch.as_tensor(f, device="cuda").requires_grad_(True)
    s = DifferentiableFESolver(m, kappa=k)
    u = s(ft)
    (u * torch.as_tensor(g, device="cuda")).sum().backward()
    uo = P.forward(nodes, el, bc, kap, f)
    gko, gfo = P.adjoint_and_grads(nodes, el, bc, kap, uo, g)
    assert relerr(u.detach().cpu().numpy(), uo) <= TOL
    gk = k.grad.cpu().numpy()
    if per_elem:
        assert np.abs(gk - gko).max() <= TOL * np.abs(gko).sum()
    else:
        assert abs(float(gk) - gko.sum()) <= TOL * np.abs(gko).sum()
    assert np.abs(ft.grad.cpu().numpy() - gfo).max() <= TOL * np.abs(gfo).max()


def test_p2_closed_form_solutions_and_convergence():
    # -u'' = 1: the exact solution is quadratic (upstream tests/test_fem.py:85-93 with P2 elements)
    m = FEMesh.line(7).to_p2()
    u = DifferentiableFESolver(m)(torch.ones(m.n_nodes, dtype=torch.float64, device="cuda")).cpu().numpy()
    x = m.nodes[:, 0].numpy()
    assert np.abs(u - x * (1 - x) / 2).max() <= 1e-13
    errs = []
    for n in (8, 16, 32):
        m = FEMesh.rectangle(n, n).to_p2()
        ue = torch.sin(np.pi * m.nodes[:, 0]) * torch.sin(np.pi * m.nodes[:, 1])
        u = DifferentiableFESolver(m)((2 * np.pi ** 2 * ue).cuda()).cpu()
        errs.append(float((u - ue).abs().max()))
    assert errs[0] / errs[1] > 6.0 and errs[1] / errs[2] > 6.0
    # the same vertices with the reference's P1 elements (and its first-order load rule): two orders less accurate
    m1 = FEMesh.rectangle(32, 32)
    ue = torch.sin(np.pi * m1.nodes[:, 0]) * torch.sin(np.pi * m1.nodes[:, 1])
    e1 = float((DifferentiableFESolver(m1)((2 * np.pi ** 2 * ue).cuda()).cpu() - ue).abs().max())
    assert errs[2] < 0.02 * e1


def test_p2_batched_and_per_sample_kappa():
    m = FEMesh.rectangle(6, 4).to_p2()
    rng = np.random.default_rng(5)
    f = rng.uniform(-1, 1, (3, m.n_nodes))
    kap = np.exp(rng.uniform(-1, 1, (3, 1)))
    s = DifferentiableFESolver(m, kappa=torch.as_tensor(kap, device="cuda"))
    u = s(torch.as_tensor(f, device="cuda")).cpu().numpy()
    for b in range(3):
        uo = P.forward(m.nodes.numpy(), m.elements.numpy(), m.dirichlet_nodes, float(kap[b, 0]), f[b])
        assert relerr(u[b], uo) <= TOL
    assert s.last_pcg is not None


def test_p2_refuses_p1_only_routes():
    m = FEMesh.rectangle(8, 8).to_p2()
    nm = m._native(torch.cuda.current_device())
    L = _native.lib()
    assert not L.dfe_band_supported(nm.handle) and not L.dfe_batch_supported(nm.handle) and not L.dfe_mg_supported(nm.handle)
    assert nm.info.chain1d == 0
