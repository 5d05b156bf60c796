"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path, called through the
reference-facing API / the C ABI, against the golden vectors of the unmodified reference and against
the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): bit-exact for CSR indices and assembled values; 1e-12 relative
(max-norm) for u / gradients in 1D; 1e-9 relative in 2D with PCG run to a 1e-13 recursive residual.
dL/dkappa is compared relative to sum_e |dL/dkappa_e| (the sum cancels).
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from difffe_physics_lab_b200 import _native
from diffhe.loss import PhysicsLoss
from diffhe.mesh import FEMesh
from diffhe.neural import NeuralPDE
from diffhe.solver import DifferentiableFESolver
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL1D = 1e-12
TOL2D = 1e-9


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _native.build()


def mesh_from_case(c):
    return FEMesh(nodes=torch.from_numpy(c["nodes"].copy()), elements=torch.from_numpy(c["elements"].copy()),
                  dirichlet_nodes=dict(c["bc"]))


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def run(mesh, kappa, f, gbar, dev="cuda", **kw):
    """u, dL/dkappa, dL/df for L = sum(gbar*u)."""
    k = torch.as_tensor(np.asarray(kappa, dtype=np.float64), device=dev).requires_grad_(True)
    ft = torch.as_tensor(np.asarray(f, dtype=np.float64), device=dev).requires_grad_(True)
    s = DifferentiableFESolver(mesh, kappa=k, **kw)
    u = s(ft)
    (u * torch.as_tensor(np.asarray(gbar, dtype=np.float64), device=dev)).sum().backward()
    return u.detach().cpu().numpy(), k.grad.cpu().numpy(), ft.grad.cpu().numpy(), s


def oracle_run(mesh, kappa, f, gbar):
    nodes, el, bc = mesh.nodes.numpy(), mesh.elements.numpy(), mesh.dirichlet_nodes
    u = O.forward(nodes, el, bc, kappa, f)
    gk, gf, _ = O.adjoint_and_grads(nodes, el, bc, kappa, u, gbar)
    return u, gk, gf


# =============================================================================== 1-D fused path
CASES_1D = ["c1_line20", "line10_bc12_rand", "line40_rand", "line8_left_only", "line8_right_only", "line2_bc",
            "line3_bc", "line100_ones", "line400_ones"]


@pytest.mark.parametrize("name", CASES_1D)
def test_1d_golden(golden, name):
    c = golden.case(name)
    m = mesh_from_case(c)
    assert m._native(torch.cuda.current_device()).info.chain1d == 1
    u, gk, gf, _ = run(m, float(c["kappa"]), c["f"], c["gbar"])
    assert relerr(u, c["u"]) <= TOL1D
    uo, gko, gfo = oracle_run(m, float(c["kappa"]), c["f"], c["gbar"])
    assert relerr(u, uo) <= TOL1D
    assert abs(float(gk) - gko.sum()) <= TOL1D * np.abs(gko).sum()
    assert abs(float(gk) - float(c["gkappa"])) <= 1e-10 * np.abs(gko).sum()      # the reference's own LU noise
    assert np.abs(gf - gfo).max() <= TOL1D * max(np.abs(gfo).max(), 1e-300)
    assert np.array_equal(u[list(c["bc"].keys())], np.array(list(c["bc"].values())))   # u[d] == g exactly


def test_1d_c1_exact_constants(golden):
    """BASELINE config 1: line(20), kappa=1, f=1."""
    c = golden.case("c1_line20")
    m = FEMesh.line(20)
    u, gk, gf, _ = run(m, 1.0, np.ones(21), np.ones(21))
    x = m.nodes[:, 0].numpy()
    assert np.abs(u - x * (1 - x) / 2).max() < 5e-16
    assert abs(float(gk) - (-1.6625)) < 1e-14
    assert np.abs(gf - c["gf"]).max() < 1e-15 and gf[0] == 0.0 and gf[-1] == 0.0
    d = golden.case("demo_line30")                      # demo step 0 (examples/poisson_1d_demo.py:104-110)
    m = FEMesh.line(30)
    with torch.no_grad():
        u_data = DifferentiableFESolver(m, kappa=torch.tensor(2.0, dtype=torch.float64))(torch.ones(31, dtype=torch.float64))
    kest = torch.tensor(1.0, dtype=torch.float64, requires_grad=True)
    u_est = DifferentiableFESolver(m, kappa=kest.abs())(torch.ones(31, dtype=torch.float64))
    loss = ((u_est - u_data) ** 2).mean()
    loss.backward()
    assert u_est.device.type == "cpu" and u_est.dtype == torch.float64      # CPU in -> CPU out
    assert abs(float(loss) - float(d["loss"])) < 1e-16
    assert abs(float(kest.grad) - float(d["gkappa"])) < 1e-15


def test_1d_reference_tests():
    """Upstream tests/test_fem.py:85-155 against the CUDA path."""
    for n, atol in ((10, 1e-10), (100, 1e-9)):
        m = FEMesh.line(n_elements=n)
        x = m.nodes.squeeze(1)
        u = DifferentiableFESolver(m)(torch.ones_like(x))
        assert torch.allclose(u, x * (1.0 - x) / 2.0, atol=atol)
        assert abs(float(u[0])) < 1e-12 and abs(float(u[-1])) < 1e-12
    errs = []
    for n in (10, 20, 40, 80):
        m = FEMesh.line(n_elements=n)
        x = m.nodes.squeeze(1)
        u = DifferentiableFESolver(m)((math.pi ** 2) * torch.sin(math.pi * x))
        errs.append(float((u - torch.sin(math.pi * x)).abs().max()))
    assert all(errs[i - 1] / (errs[i] + 1e-15) > 3.0 for i in range(1, 4))
    m = FEMesh.line(n_elements=10, bc_left=1.0, bc_right=2.0)
    x = m.nodes.squeeze(1)
    assert torch.allclose(DifferentiableFESolver(m)(torch.zeros_like(x)), 1.0 + x, atol=1e-10)
    kappa = torch.tensor(1.0, dtype=torch.float64, requires_grad=True)
    m = FEMesh.line(n_elements=5)
    DifferentiableFESolver(m, kappa=kappa)(torch.ones(6, dtype=torch.float64)).sum().backward()
    assert kappa.grad is not None and kappa.grad.abs() > 1e-10
    f32 = torch.linspace(0, 1, 13, dtype=torch.float32)
    assert DifferentiableFESolver(FEMesh.line(12))(f32).dtype == torch.float64


@pytest.mark.parametrize("n,B,bcs", [(4351, 3, (0.0, 0.0)), (4352, 2, (0.3, -0.2)), (5000, 4, (0.5, None)),
                                     (20000, 3, (None, 1.5)), (100000, 3, (0.0, 0.0))])
def test_1d_multi_chunk_vs_exact_oracle(n, B, bcs):
    """Sizes spanning 1, 2, 5 and 23 chunks per sample; per-sample kappa; random f and gbar; the oracle
    solves the same float64 system in 50-digit arithmetic."""
    rng = np.random.default_rng(n)
    m = FEMesh.line(n, x_left=-0.2, x_right=1.3, bc_left=bcs[0], bc_right=bcs[1])
    f = rng.uniform(0, 1, (B, n + 1))
    gbar = rng.standard_normal((B, n + 1))
    kap = np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1)))
    u, gk, gf, _ = run(m, kap, f, gbar)
    assert u.shape == (B, n + 1) and gk.shape == (B, 1) and gf.shape == (B, n + 1)
    for b in range(B):
        uo, gko, gfo = oracle_run(m, float(kap[b, 0]), f[b], gbar[b])
        assert relerr(u[b], uo) <= TOL1D, f"sample {b}"
        assert abs(gk[b, 0] - gko.sum()) <= TOL1D * np.abs(gko).sum()
        assert np.abs(gf[b] - gfo).max() <= TOL1D * np.abs(gfo).max()


def test_1d_shared_scalar_kappa_batch_and_determinism():
    rng = np.random.default_rng(5)
    n, B = 9000, 6
    m = FEMesh.line(n)
    f = rng.uniform(0.5, 1.5, (B, n + 1))
    gbar = rng.standard_normal((B, n + 1))
    u, gk, gf, _ = run(m, 1.37, f, gbar)
    u2, gk2, gf2, _ = run(m, 1.37, f, gbar)
    assert np.array_equal(u, u2) and np.array_equal(gk, gk2) and np.array_equal(gf, gf2)   # no float atomics
    tot, mag = 0.0, 0.0
    for b in range(B):
        uo, gko, _ = oracle_run(m, 1.37, f[b], gbar[b])
        assert relerr(u[b], uo) <= TOL1D
        tot += gko.sum()
        mag += np.abs(gko).sum()
    assert gk.shape == () and abs(float(gk) - tot) <= TOL1D * mag
    # the un-batched call returns (n_nodes,) and equals row 0
    u0, _, _, _ = run(m, 1.37, f[0], gbar[0])
    assert u0.shape == (n + 1,) and np.array_equal(u0, u[0])


@pytest.mark.parametrize("n,bcs", [(40, (0.0, 0.0)), (2304, (0.3, -0.2)), (6001, (0.5, None)), (30000, (None, 1.5))])
def test_1d_per_element_kappa(n, bcs):
    """kappa of shape (n_el,) (shared field) and (B, n_el) on the fused 1-D path (reference loop with
    kappa -> kappa[e]), against the exact oracle."""
    rng = np.random.default_rng(n + 7)
    B = 3
    m = FEMesh.line(n, x_left=0.1, x_right=1.9, bc_left=bcs[0], bc_right=bcs[1])
    assert m._native(torch.cuda.current_device()).info.chain1d == 1
    f = rng.uniform(0, 1, (B, n + 1))
    gbar = rng.standard_normal((B, n + 1))
    kap = np.exp(rng.uniform(np.log(0.1), np.log(10.0), (B, n)))
    u, gk, gf, _ = run(m, kap, f, gbar)                       # per-sample per-element field
    assert gk.shape == (B, n)
    nodes, el, bc = m.nodes.numpy(), m.elements.numpy(), m.dirichlet_nodes
    for b in range(B):
        uo, _, gfo = oracle_run(m, kap[b], f[b], gbar[b])
        assert relerr(u[b], uo) <= TOL1D
        # dL/dkappa_e = -(dlambda_e)(du_e)/h_e differences neighbouring u values: it is a function of the STORED
        # float64 u (ulp(u)/|du_e| ~ 1e-11 here), so the oracle differentiates the same stored u as the kernel
        gko, _, _ = O.adjoint_and_grads(nodes, el, bc, kap[b], u[b], gbar[b])
        assert np.abs(gk[b] - gko).max() <= TOL1D * np.abs(gko).max()
        assert np.abs(gf[b] - gfo).max() <= TOL1D * np.abs(gfo).max()
    u, gk, gf, _ = run(m, kap[0], f, gbar)                    # one field shared by the batch
    assert gk.shape == (n,)
    tot = np.zeros(n)
    for b in range(B):
        uo, _, gfo = oracle_run(m, kap[0], f[b], gbar[b])
        assert relerr(u[b], uo) <= TOL1D
        assert np.abs(gf[b] - gfo).max() <= TOL1D * np.abs(gfo).max()
        gko, _, _ = O.adjoint_and_grads(nodes, el, bc, kap[0], u[b], gbar[b])
        tot += gko
    assert np.abs(gk - tot).max() <= TOL1D * np.abs(tot).max() * B
    u2, gk2, gf2, _ = run(m, kap[0], f, gbar)
    assert np.array_equal(u, u2) and np.array_equal(gk, gk2) and np.array_equal(gf, gf2)


def test_1d_unaligned_views_and_strides():
    """Rows that start on 8-byte (not 16-byte) boundaries and non-unit batch strides."""
    rng = np.random.default_rng(9)
    n, B = 6001, 3                      # n_nodes = 6002 (even) — shift by a storage offset instead
    m = FEMesh.line(n, bc_left=0.1, bc_right=0.2)
    big = torch.tensor(rng.uniform(0, 1, (B, n + 1 + 5)), device="cuda")
    f = big[:, 3:3 + n + 1]            # storage offset 3 doubles, row stride n+6
    assert not f.is_contiguous()
    s = DifferentiableFESolver(m, kappa=0.9)
    u_view = s(f)
    u_cont = s(f.contiguous())
    assert torch.equal(u_view, u_cont)   # the host layer makes the view contiguous: same kernel, same bits
    L = _native.lib()                   # straight through the C ABI with ld > n_nodes and odd offsets
    nm = m._native(torch.cuda.current_device())
    out = torch.zeros((B, n + 1 + 3), dtype=torch.float64, device="cuda")
    kap = torch.tensor([0.9], dtype=torch.float64, device="cuda")
    ws = torch.empty(L.dfe_solve1d_workspace_bytes(nm.handle, B), dtype=torch.uint8, device="cuda")
    rc = L.dfe_solve1d_fwd(nm.handle, B, f.data_ptr(), f.stride(0), kap.data_ptr(), 0, -1,
                           out[:, 1:].data_ptr(), out.stride(0), ws.data_ptr(), ws.numel(),
                           torch.cuda.current_stream().cuda_stream)
    assert rc == 0, L.dfe_last_error()
    torch.cuda.synchronize()
    # rows on odd 8-byte offsets take the two-pass kernels (different summation order than the pipelined one)
    assert float((out[:, 1:n + 2] - u_cont).abs().max()) <= 1e-13 * float(u_cont.abs().max())
    assert float(out[:, 0].abs().max()) == 0.0 and float(out[:, n + 2:].abs().max()) == 0.0


def test_1d_full_size_properties():
    """BASELINE config 2 shape (n_el = 100000) with a small batch: size-independent properties."""
    rng = np.random.default_rng(11)
    n, B = 100000, 4
    m = FEMesh.line(n)
    f = torch.tensor(rng.uniform(0, 1, (B, n + 1)), device="cuda")
    kap = torch.tensor(np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1))), device="cuda")
    s = DifferentiableFESolver(m, kappa=kap)
    u = s(f)
    # linearity in f
    a, b = 0.75, -1.5
    f_mix = (a * f[0] + b * f[1]).expand(B, -1).contiguous()
    k_same = kap[0:1].expand(B, 1).contiguous()
    u_mix = DifferentiableFESolver(m, kappa=k_same)(f_mix)
    u_a = DifferentiableFESolver(m, kappa=k_same)(f[0:1].expand(B, -1).contiguous())
    u_b = DifferentiableFESolver(m, kappa=k_same)(f[1:2].expand(B, -1).contiguous())
    assert float((u_mix - (a * u_a + b * u_b)).abs().max() / u_mix.abs().max()) <= 1e-12
    # u scales as 1/kappa only up to the rounding of the assembled diagonals fl(k_{i-1}+k_i): the float64
    # systems for different kappa are not proportional (|M^-1 E| ~ 1e-7 here) — the very effect the
    # Neumann sweep resolves.  So this identity is a 1e-6 check, not a parity check.
    u_k1 = DifferentiableFESolver(m, kappa=torch.ones(B, 1, dtype=torch.float64, device="cuda"))(f)
    assert float((u * kap - u_k1).abs().max() / u_k1.abs().max()) <= 1e-6
    # adjoint consistency: <gbar, u(f)> == <dL/df, f>  (u is linear in f for zero Dirichlet data)
    gbar = torch.tensor(rng.standard_normal((B, n + 1)), device="cuda")
    fr = f.clone().requires_grad_(True)
    kr = kap.clone().requires_grad_(True)
    ur = DifferentiableFESolver(m, kappa=kr)(fr)
    (ur * gbar).sum().backward()
    lhs = (gbar * ur.detach()).sum(dim=1)
    rhs = (fr.grad * f).sum(dim=1)
    assert float(((lhs - rhs).abs() / (gbar * ur.detach()).abs().sum(dim=1)).max()) <= 1e-12
    # scalar kappa, zero BC: dL/dkappa = -L/kappa (same 1e-6 caveat)
    assert float(((kr.grad[:, 0] + lhs / kap[:, 0]).abs() / (gbar * ur.detach()).abs().sum(dim=1) * kap[:, 0]).max()) <= 1e-6


def test_1d_kappa_recovery_demo():
    """examples/poisson_1d_demo.py:88-112: 200 Adam steps through the solver recover kappa = 2.0000."""
    m = FEMesh.line(n_elements=30)
    f = torch.ones(m.n_nodes, dtype=torch.float64)
    with torch.no_grad():
        u_data = DifferentiableFESolver(m, kappa=torch.tensor(2.0, dtype=torch.float64))(f)
    k = torch.tensor(1.0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([k], lr=0.1)
    for _ in range(200):
        opt.zero_grad()
        loss = ((DifferentiableFESolver(m, kappa=k.abs())(f) - u_data) ** 2).mean()
        loss.backward()
        opt.step()
    assert f"{float(k.abs()):.4f}" == "2.0000"


# =============================================================================== general / 2-D path
def abi_assemble(m, kappa, f):
    """dfe_assemble + dfe_eliminate straight through the C ABI; returns host arrays."""
    L = _native.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    nm = m._native(dev.index)
    I = nm.info
    kap = torch.as_tensor(np.atleast_1d(np.asarray(kappa, dtype=np.float64)), device=dev)
    mode = _native.KAPPA_SCALAR if kap.numel() == 1 else _native.KAPPA_PER_ELEMENT
    ft = torch.as_tensor(np.asarray(f, dtype=np.float64), device=dev)
    vals = torch.empty(I.nnz_full, dtype=torch.float64, device=dev)
    F = torch.empty(I.n_nodes, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _native.check(L.dfe_assemble(nm.handle, kap.data_ptr(), mode, ft.data_ptr(), vals.data_ptr(), F.data_ptr(), st))
    vf = torch.empty(max(I.nnz_free, 1), dtype=torch.float64, device=dev)
    sell = torch.empty(max(I.sell_nnz, 1), dtype=torch.float64, device=dev)
    Ff = torch.empty(max(I.n_free, 1), dtype=torch.float64, device=dev)
    dinv = torch.empty(max(I.n_free, 1), dtype=torch.float64, device=dev)
    _native.check(L.dfe_eliminate(nm.handle, vals.data_ptr(), F.data_ptr(), vf.data_ptr(), sell.data_ptr(), Ff.data_ptr(), dinv.data_ptr(), st))
    torch.cuda.synchronize()
    return nm, vals.cpu().numpy(), F.cpu().numpy(), vf.cpu().numpy()[:I.nnz_free], Ff.cpu().numpy()[:I.n_free], dinv.cpu().numpy()[:I.n_free]


CASES_2D = ["rect4_ones", "rect8_ones", "rect16_ones", "rect7x5_rand", "rect6x9_rand", "rect5x4_partial_bc", "rect3_degenerate"]
CASES_DENSE = ["rect4_ones", "rect8_ones", "rect7x5_rand", "rect6x9_rand", "rect5x4_partial_bc", "rect3_degenerate",
               "c1_line20", "line40_rand", "line8_right_only", "line2_bc"]


@pytest.mark.parametrize("name", CASES_DENSE)
def test_assembly_bit_exact_vs_reference_dense(golden, name):
    """CSR indices and values from dfe_assemble equal the reference's dense K, F bit for bit."""
    c = golden.case(name)
    m = mesh_from_case(c)
    n = m.n_nodes
    nm, vals, F, vf, Ff, dinv = abi_assemble(m, float(c["kappa"]), c["f"])
    rp, col = nm.csr(0)
    dense = np.zeros((n, n))
    dense[np.repeat(np.arange(n), np.diff(rp)), col] = vals
    assert np.array_equal(dense, c["K"]) and np.array_equal(F, c["F"])
    orp, ocol, ovals, oF = O.assemble_csr(c["nodes"], c["elements"], float(c["kappa"]), c["f"])
    free, frp, fcol, fvals, oFf = O.apply_bc(orp, ocol, ovals, oF, c["bc"])
    rpf, colf = nm.csr(1)
    assert np.array_equal(rpf, frp) and np.array_equal(colf, fcol)
    assert np.array_equal(vf, fvals) and np.array_equal(Ff, oFf)
    diag = fvals[[int(np.nonzero(fcol[frp[r]:frp[r + 1]] == r)[0][0]) + frp[r] for r in range(len(free))]]
    assert np.array_equal(dinv, 1.0 / diag)


def test_assembly_bit_exact_large_per_element_kappa():
    rng = np.random.default_rng(3)
    m = FEMesh.rectangle(61, 47, x_range=(0.0, 2.3), y_range=(-1.0, 0.7), bc_value=0.4)
    kap = np.exp(rng.uniform(np.log(1e-3), 0.0, m.n_elements))
    f = rng.uniform(-1, 1, m.n_nodes)
    nm, vals, F, vf, Ff, _ = abi_assemble(m, kap, f)
    orp, ocol, ovals, oF = O.assemble_csr(m.nodes.numpy(), m.elements.numpy(), kap, f)
    _, frp, fcol, fvals, oFf = O.apply_bc(orp, ocol, ovals, oF, m.dirichlet_nodes)
    assert np.array_equal(vals, ovals) and np.array_equal(F, oF)
    assert np.array_equal(vf, fvals) and np.array_equal(Ff, oFf)
    assert np.array_equal(np.sort(vals), np.sort(ovals))


ASM_ENVS = ("DFE_ASSEMBLE_ROWS", "DFE_ASSEMBLE_GENERAL")


@pytest.mark.parametrize("nx,ny,scalar,scale", [(300, 70, False, 1.0), (31, 7, True, 1.0), (32, 8, False, 1.0), (1, 1, False, 1.0),
                                                (95, 130, True, 1.0), (64, 200, False, 1.0), (40, 33, False, 1e130),
                                                (33, 21, False, 1e-9)])
def test_assembly_three_kernels_same_bits(monkeypatch, nx, ny, scalar, scale):
    """The element-parallel tile kernel (default on rectangle() patterns), the row-owner structured kernel and the general
    adjacency-list kernel produce the same K and F bit for bit — tile edges, mesh edges, partial Dirichlet sets, coordinates
    outside the range in which the handle lets the tile kernel skip its per-numerator range checks (scale 1e130), a mesh
    whose triangles are all below the reference's degenerate-area threshold (scale 1e-9) — and, on the small cases, the
    oracle's reference-order accumulation."""
    rng = np.random.default_rng(nx * 1000 + ny)
    m = FEMesh.rectangle(nx, ny, x_range=(-0.7 * scale, 1.9 * scale), y_range=(0.1 * scale, 0.9 * scale), bc_value=0.25)
    if nx > 4:                                   # a partial Dirichlet set keeps the structured pattern
        for k in list(m.dirichlet_nodes)[::3]:
            del m.dirichlet_nodes[k]
    kap = 1.7 if scalar else np.exp(rng.uniform(np.log(1e-3), 0.0, m.n_elements))
    f = rng.uniform(-1, 1, m.n_nodes)
    out = {}
    for name, env in (("tile", None), ("rows", "DFE_ASSEMBLE_ROWS"), ("general", "DFE_ASSEMBLE_GENERAL")):
        for e in ASM_ENVS:
            monkeypatch.delenv(e, raising=False)
        if env:
            monkeypatch.setenv(env, "1")
        _, vals, F, vf, Ff, dinv = abi_assemble(m, kap, f)
        out[name] = (vals, F, vf, Ff, dinv)
    for name in ("rows", "general"):
        for a, b in zip(out["tile"], out[name]):
            assert np.array_equal(a, b), name
    if m.n_nodes < 30000:
        orp, ocol, ovals, oF = O.assemble_csr(m.nodes.numpy(), m.elements.numpy(), kap, f)
        assert np.array_equal(out["tile"][0], ovals) and np.array_equal(out["tile"][1], oF)


@pytest.mark.parametrize("name", ["rect7x5_rand", "rect5x4_partial_bc", "line40_rand"])
def test_assemble_sparse_public_api(golden, name):
    """``assemble_sparse`` (the reference roadmap's sparse assembly): the torch CSR matrix densifies to the reference's K bit for
    bit, F likewise; the matrix applies like the dense one."""
    import warnings
    from difffe_physics_lab_b200 import assemble_sparse
    c = golden.case(name)
    m = mesh_from_case(c)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        K, F = assemble_sparse(m, float(c["kappa"]), torch.as_tensor(c["f"]))
        assert K.is_cuda and K.layout == torch.sparse_csr
        assert np.array_equal(K.to_dense().cpu().numpy(), c["K"]) and np.array_equal(F.cpu().numpy(), c["F"])
        x = torch.linspace(0, 1, m.n_nodes, dtype=torch.float64, device="cuda")
        assert np.allclose((K @ x).cpu().numpy(), c["K"] @ x.cpu().numpy(), rtol=0, atol=1e-12 * np.abs(c["K"]).max())


@pytest.mark.parametrize("name", CASES_2D)
def test_2d_golden(golden, name):
    c = golden.case(name)
    m = mesh_from_case(c)
    u, gk, gf, s = run(m, float(c["kappa"]), c["f"], c["gbar"])
    assert relerr(u, c["u"]) <= TOL2D
    uo, gko, gfo = oracle_run(m, float(c["kappa"]), c["f"], c["gbar"])
    assert relerr(u, uo) <= TOL2D
    assert abs(float(gk) - gko.sum()) <= TOL2D * np.abs(gko).sum()
    assert abs(float(gk) - float(c["gkappa"])) <= TOL2D * np.abs(gko).sum()
    assert np.abs(gf - gfo).max() <= TOL2D * np.abs(gfo).max()
    assert np.abs(gf - c["gf"]).max() <= TOL2D * np.abs(c["gf"]).max()


def test_2d_reference_tests():
    """Upstream tests/test_fem.py:163-179."""
    m = FEMesh.rectangle(nx=4, ny=4)
    assert DifferentiableFESolver(m)(torch.zeros(m.n_nodes, dtype=torch.float64)).abs().max() < 1e-10
    m = FEMesh.rectangle(nx=8, ny=8)
    u = DifferentiableFESolver(m)(torch.ones(m.n_nodes, dtype=torch.float64))
    assert float(u[m.free_nodes()].min()) > 0.0


def test_2d_c3_vs_literal_reference(golden, golden_big):
    """BASELINE config 3: 128x128, f=1 — against the literal dense reference solve (147 s on CPU)."""
    for n, iters in ((32, 72), (64, None), (128, 296)):
        ref = golden.case("rect32_ones")["u"] if n == 32 else golden_big.case(f"rect{n}_ones")["u"]
        m = FEMesh.rectangle(n, n)
        k = torch.tensor(1.0, dtype=torch.float64, device="cuda", requires_grad=True)
        sj = DifferentiableFESolver(m, kappa=k, solver2d="jacobi")                # north_star's SpMV + Jacobi-PCG
        uj = sj(torch.ones(m.n_nodes, dtype=torch.float64, device="cuda"))
        assert relerr(uj.detach().cpu().numpy(), ref) <= TOL2D
        if iters:
            assert abs(sj.last_pcg[0][0] - iters) <= 3         # SURVEY §7 hard part 2
        s = DifferentiableFESolver(m, kappa=k)                                    # default: multigrid-preconditioned CG
        u = s(torch.ones(m.n_nodes, dtype=torch.float64, device="cuda"))
        assert relerr(u.detach().cpu().numpy(), ref) <= TOL2D
        assert s._opts["last_solver2d"] == ("mg" if n >= 34 else "jacobi")       # 31 x 31 unknowns: too small to bother
        assert s._opts["last_solver2d"] == "jacobi" or s.last_pcg[0][0] <= 20
        u.sum().backward()
        # scalar kappa, zero BC: d(sum u)/dkappa = -sum u / kappa
        assert abs(float(k.grad) + float(u.sum())) <= TOL2D * abs(float(u.sum()))
    assert abs(float(u.max()) - 0.07366781046909493) < 1e-11 and abs(float(u.sum()) - 575.6892139032584) < 1e-6


def test_2d_per_element_kappa():
    """SURVEY §8(c) extension oracle + a random heterogeneous field (config 4 in miniature)."""
    m = FEMesh.rectangle(4, 4)
    kap = torch.tensor(1.0 + 0.5 * np.arange(32) / 32, device="cuda", requires_grad=True)
    f = torch.ones(25, dtype=torch.float64, device="cuda")
    u = DifferentiableFESolver(m, kappa=kap)(f)
    _, _, _, F = O.assemble_csr(m.nodes.numpy(), m.elements.numpy(), kap.detach().cpu().numpy(), np.ones(25))
    J = (torch.tensor(F, device="cuda") * u).sum()
    J.backward()
    assert abs(float(J) - 0.02336457495061802) < 1e-12
    ref8 = [0, -0.00147538203336745, -0.00073769101668373, -0.00118563863606154, -0.00114170043120912,
            -0.00073796675704933, -0.00067784987673196, -0.00067784987673196]
    assert np.abs(kap.grad[:8].cpu().numpy() - ref8).max() < 1e-12
    assert abs(float(kap.grad.sum()) + 0.019090990124307906) < 1e-12
    rng = np.random.default_rng(4)
    m = FEMesh.rectangle(40, 33, x_range=(0.0, 1.7), y_range=(0.0, 1.1), bc_value=-0.3)
    kap = np.exp(rng.uniform(np.log(1e-3), 0.0, m.n_elements))
    f = rng.uniform(0, 1, m.n_nodes)
    gbar = rng.standard_normal(m.n_nodes)
    u, gk, gf, _ = run(m, kap, f, gbar)
    uo, gko, gfo = oracle_run(m, kap, f, gbar)
    assert relerr(u, uo) <= TOL2D
    assert np.abs(gk - gko).max() <= TOL2D * np.abs(gko).max()
    assert np.abs(gf - gfo).max() <= TOL2D * np.abs(gfo).max()


def test_2d_batched_and_determinism():
    rng = np.random.default_rng(6)
    m = FEMesh.rectangle(20, 15, bc_value=0.1)
    B = 3
    f = rng.uniform(0, 1, (B, m.n_nodes))
    gbar = rng.standard_normal((B, m.n_nodes))
    u, gk, gf, _ = run(m, 1.3, f, gbar)
    u2, gk2, gf2, _ = run(m, 1.3, f, gbar)
    assert np.array_equal(u, u2) and np.array_equal(gk, gk2) and np.array_equal(gf, gf2)
    tot, mag = 0.0, 0.0
    for b in range(B):
        uo, gko, gfo = oracle_run(m, 1.3, f[b], gbar[b])
        assert relerr(u[b], uo) <= TOL2D and np.abs(gf[b] - gfo).max() <= TOL2D * np.abs(gfo).max()
        tot += gko.sum()
        mag += np.abs(gko).sum()
    assert abs(float(gk) - tot) <= TOL2D * mag
    kap = np.array([[0.5], [1.0], [2.0]])
    u, gk, _, _ = run(m, kap, f, gbar)
    for b in range(B):
        uo, gko, _ = oracle_run(m, float(kap[b, 0]), f[b], gbar[b])
        assert relerr(u[b], uo) <= TOL2D and abs(gk[b, 0] - gko.sum()) <= TOL2D * np.abs(gko).sum()


def test_2d_batched_small_systems_kernel():
    """dfe_batch_fwd / dfe_batch_bwd (one CTA per sample, shared matrix): against the oracle, against the per-sample
    route (dfe_assemble / dfe_eliminate / dfe_pcg per system), more samples than CTAs, per-element shared kappa,
    the largest mesh the kernel takes, and a 1-D mesh that is not a chain."""
    L = _native.lib()
    rng = np.random.default_rng(11)
    # (a) more samples than resident CTAs, non-zero Dirichlet data, scalar shared kappa
    m = FEMesh.rectangle(9, 7, x_range=(0.0, 1.3), y_range=(-0.4, 0.5), bc_value=0.3)
    assert L.dfe_batch_supported(m._native(torch.cuda.current_device()).handle) == 1
    B = 700
    f = rng.uniform(0, 1, (B, m.n_nodes))
    gbar = rng.standard_normal((B, m.n_nodes))
    assert L.dfe_band_supported(m._native(torch.cuda.current_device()).handle) == 1
    for route in ("band", "pcg"):            # banded Cholesky (default for this mesh) and one-CTA-per-sample PCG
        k = torch.tensor(0.8, dtype=torch.float64, device="cuda", requires_grad=True)
        ft = torch.tensor(f, device="cuda", requires_grad=True)
        s = DifferentiableFESolver(m, kappa=k)
        s._opts["batch_solver"] = route
        ut = s(ft)
        (ut * torch.tensor(gbar, device="cuda")).sum().backward()
        assert (s.last_pcg[0][0] > 0) == (route == "pcg")
        u, gf = ut.detach().cpu().numpy(), ft.grad.cpu().numpy()
        tot, mag = 0.0, 0.0
        for b in list(range(0, B, 97)) + [B - 1]:
            uo, gko, gfo = oracle_run(m, 0.8, f[b], gbar[b])
            assert relerr(u[b], uo) <= TOL2D and np.abs(gf[b] - gfo).max() <= TOL2D * np.abs(gfo).max()
        if route == "band":
            u_band, gk_band = u, float(k.grad)
        else:
            assert relerr(u, u_band) <= 1e-11 and abs(float(k.grad) - gk_band) <= 1e-10 * abs(gk_band)
    # (a') the stencil-form load / gradient kernels with 1, 2 and 4 samples per CTA barrier: same bits
    import os
    outs = []
    for spi in ("1", "2", "4"):
        os.environ["DFE_BAND_SPI"] = spi
        try:
            k = torch.tensor(0.8, dtype=torch.float64, device="cuda", requires_grad=True)
            ft = torch.tensor(f, device="cuda", requires_grad=True)
            s = DifferentiableFESolver(m, kappa=k)
            ut = s(ft)
            (ut * torch.tensor(gbar, device="cuda")).sum().backward()
            outs.append((ut.detach().clone(), ft.grad.clone(), k.grad.clone()))
        finally:
            del os.environ["DFE_BAND_SPI"]
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(outs[0], o))
    assert torch.equal(outs[2][0], torch.tensor(u_band, device="cuda"))
    # (b) batched route == per-sample route (same arithmetic for F, lifting, SpMV; different reduction trees)
    B = 6
    k = torch.tensor(1.7, dtype=torch.float64, device="cuda", requires_grad=True)
    ft = torch.tensor(f[:B], device="cuda", requires_grad=True)
    s1 = DifferentiableFESolver(m, kappa=k)
    u1 = s1(ft)
    (u1 * torch.tensor(gbar[:B], device="cuda")).sum().backward()
    gk1, gf1 = k.grad.clone(), ft.grad.clone()
    k.grad = None
    ft.grad = None
    s2 = DifferentiableFESolver(m, kappa=k)
    s2._opts["batch_min"] = 10 ** 9          # force one cooperative PCG per sample
    u2 = s2(ft)
    (u2 * torch.tensor(gbar[:B], device="cuda")).sum().backward()
    assert float((u1 - u2).abs().max()) <= 1e-12 * float(u2.abs().max())
    assert abs(float(gk1 - k.grad)) <= 1e-11 * abs(float(k.grad)) + 1e-14
    assert float((gf1 - ft.grad).abs().max()) <= 1e-12 * float(ft.grad.abs().max())
    # (c) shared per-element field
    m = FEMesh.rectangle(13, 11, bc_value=-0.2)
    kap = np.exp(rng.uniform(np.log(1e-2), 0.0, m.n_elements))
    B = 4
    f = rng.uniform(0, 1, (B, m.n_nodes))
    gbar = rng.standard_normal((B, m.n_nodes))
    u, gk, gf, _ = run(m, kap, f, gbar)
    gsum = np.zeros(m.n_elements)
    for b in range(B):
        uo, gko, gfo = oracle_run(m, kap, f[b], gbar[b])
        assert relerr(u[b], uo) <= TOL2D and np.abs(gf[b] - gfo).max() <= TOL2D * np.abs(gfo).max()
        gsum += gko
    assert np.abs(gk - gsum).max() <= TOL2D * np.abs(gsum).max()
    # (d) largest mesh the shared-memory PCG kernel takes (bandwidth 46 > 32: no banded route), and one beyond it
    m = FEMesh.rectangle(45, 44)
    assert m.n_nodes - len(m.dirichlet_nodes) == 44 * 43
    assert L.dfe_batch_supported(m._native(torch.cuda.current_device()).handle) == 1
    assert L.dfe_band_supported(m._native(torch.cuda.current_device()).handle) == 0
    f = rng.uniform(0, 1, (3, m.n_nodes))
    gbar = rng.standard_normal((3, m.n_nodes))
    u, gk, gf, _ = run(m, 1.1, f, gbar)
    uo, gko, gfo = oracle_run(m, 1.1, f[2], gbar[2])
    assert relerr(u[2], uo) <= TOL2D and np.abs(gf[2] - gfo).max() <= TOL2D * np.abs(gfo).max()
    assert L.dfe_batch_supported(FEMesh.rectangle(60, 60)._native(torch.cuda.current_device()).handle) == 0
    # long thin mesh: too many unknowns for the shared-memory kernel, but banded (half bandwidth 9): direct route
    m = FEMesh.rectangle(8, 400, y_range=(0.0, 30.0), bc_value=0.05)
    assert L.dfe_batch_supported(m._native(torch.cuda.current_device()).handle) == 0
    assert L.dfe_band_supported(m._native(torch.cuda.current_device()).handle) == 1
    f = rng.uniform(0, 1, (5, m.n_nodes))
    gbar = rng.standard_normal((5, m.n_nodes))
    u, gk, gf, _ = run(m, 0.6, f, gbar)
    tot, mag = 0.0, 0.0
    for b in range(5):
        uo, gko, gfo = oracle_run(m, 0.6, f[b], gbar[b])
        assert relerr(u[b], uo) <= TOL2D and np.abs(gf[b] - gfo).max() <= TOL2D * np.abs(gfo).max()
        tot += gko.sum()
        mag += np.abs(gko).sum()
    assert abs(float(gk) - tot) <= TOL2D * mag
    # (e) 1-D mesh that is not a chain (interior Dirichlet node): general path, batched
    m = FEMesh.line(50, bc_left=0.1, bc_right=0.4)
    m.dirichlet_nodes[20] = -0.3
    f = rng.uniform(0, 1, (5, 51))
    gbar = rng.standard_normal((5, 51))
    u, gk, gf, _ = run(m, 2.2, f, gbar)
    tot, mag = 0.0, 0.0
    for b in range(5):
        uo, gko, gfo = oracle_run(m, 2.2, f[b], gbar[b])
        assert relerr(u[b], uo) <= TOL2D and np.abs(gf[b] - gfo).max() <= TOL2D * max(np.abs(gfo).max(), 1e-300)
        tot += gko.sum()
        mag += np.abs(gko).sum()
    assert abs(float(gk) - tot) <= TOL2D * mag


def test_general_path_for_non_chain_1d_meshes():
    """1-D meshes the fused kernel does not take: interior Dirichlet node, permuted elements, per-element kappa."""
    rng = np.random.default_rng(8)
    m = FEMesh.line(60, bc_left=0.2, bc_right=-0.1)
    m.dirichlet_nodes[17] = 0.4
    perm = torch.tensor(rng.permutation(60))
    m.elements = m.elements[perm]
    assert m._native(torch.cuda.current_device()).info.chain1d == 0
    f = rng.uniform(0, 1, 61)
    gbar = rng.standard_normal(61)
    u, gk, gf, _ = run(m, 1.9, f, gbar)
    uo, gko, gfo = oracle_run(m, 1.9, f, gbar)
    assert relerr(u, uo) <= TOL2D and abs(float(gk) - gko.sum()) <= TOL2D * np.abs(gko).sum()
    assert np.abs(gf - gfo).max() <= TOL2D * np.abs(gfo).max()
    m = FEMesh.line(50)
    kap = np.exp(rng.uniform(-1, 1, 50))
    f = rng.uniform(0, 1, 51)
    gbar = rng.standard_normal(51)
    u, gk, gf, _ = run(m, kap, f, gbar)
    uo, gko, gfo = oracle_run(m, kap, f, gbar)
    assert relerr(u, uo) <= TOL2D and np.abs(gk - gko).max() <= TOL2D * np.abs(gko).max()


def test_pcg_error_reporting():
    m = FEMesh.rectangle(24, 24)
    f = torch.ones(m.n_nodes, dtype=torch.float64, device="cuda")
    with pytest.raises(_native.NotConvergedError):
        DifferentiableFESolver(m, pcg_maxit=3)(f)
    with pytest.raises(_native.NotConvergedError):          # the batched one-CTA-per-sample kernel reports it too
        sb = DifferentiableFESolver(m, pcg_maxit=3)
        sb._opts["batch_solver"] = "pcg"
        sb(torch.ones((4, m.n_nodes), dtype=torch.float64, device="cuda"))
    m = FEMesh.rectangle(6, 6)
    m.dirichlet_nodes = {}
    with pytest.raises(_native.DfeError):                   # banded Cholesky: zero pivot of the singular K
        DfeS = DifferentiableFESolver(m)
        DfeS(torch.ones((3, 49), dtype=torch.float64, device="cuda"))
    m = FEMesh.rectangle(6, 6)
    m.dirichlet_nodes = {}              # singular K: the reference silently returns garbage
    with pytest.raises(_native.DfeError):
        DifferentiableFESolver(m, pcg_maxit=200)(torch.ones(49, dtype=torch.float64, device="cuda"))


def test_2d_large_properties():
    """512x512 heterogeneous kappa (config 4 shape, reduced): residual, symmetry and compliance identities."""
    rng = np.random.default_rng(12)
    m = FEMesh.rectangle(512, 512)
    kap = torch.tensor(np.exp(rng.uniform(np.log(1e-2), 0.0, m.n_elements)), device="cuda", requires_grad=True)
    f = torch.ones(m.n_nodes, dtype=torch.float64, device="cuda")
    s = DifferentiableFESolver(m, kappa=kap)
    u = s(f)
    nm, vals, F, vf, Ff, dinv = abi_assemble(m, kap.detach().cpu().numpy(), np.ones(m.n_nodes))
    import scipy.sparse as sp
    rpf, colf = nm.csr(1)
    A = sp.csr_matrix((vf, colf, rpf))
    free = nm.free_nodes()
    uf = u.detach().cpu().numpy()[free]
    assert np.linalg.norm(A @ uf - Ff) <= 1e-9 * np.linalg.norm(Ff)           # true residual (stagnates ~1e-11)
    assert abs(A - A.T).max() == 0.0                                           # K_free bitwise symmetric
    # compliance J = F^T u: gbar = F  =>  lambda = u and dJ/dkappa_e = -u_e^T K_e^0 u_e <= 0
    J = (torch.tensor(F, device="cuda") * u).sum()
    J.backward()
    assert float(kap.grad.max()) <= 0.0
    # Euler identity for a degree -1 homogeneous function of kappa: sum_e kappa_e dJ/dkappa_e = -J
    assert abs(float((kap.detach() * kap.grad).sum()) + float(J)) <= 1e-8 * abs(float(J))


def test_neural_pde_reference_tests():
    """Upstream tests/test_neural.py:29-90 (training loops that call the hot path through PhysicsLoss)."""
    m = FEMesh.line(n_elements=10)
    model = NeuralPDE(m, hidden_dim=16, n_layers=2)
    model.train_pde(lambda x: torch.ones_like(x), n_epochs=100, verbose=False)
    u = model()
    assert abs(float(u[0])) < 1e-10 and abs(float(u[-1])) < 1e-10
    m = FEMesh.line(n_elements=20)
    model = NeuralPDE(m, hidden_dim=32, n_layers=3)
    losses = model.train_pde(lambda x: torch.ones_like(x), n_epochs=300, verbose=False)
    assert losses[-1] < losses[0]
    torch.manual_seed(42)
    model = NeuralPDE(m, hidden_dim=64, n_layers=3)
    u_fem = DifferentiableFESolver(m)(torch.ones(21, dtype=torch.float64)).detach()
    model.train_pde(lambda x: torch.ones_like(x), n_epochs=3000, lr=1e-3, verbose=False)
    free = m.free_nodes()
    rel = (model().detach()[free] - u_fem[free]).abs().max() / u_fem[free].abs().max()
    assert float(rel) < 0.05
    torch.manual_seed(0)
    m = FEMesh.line(n_elements=10)
    model = NeuralPDE(m, hidden_dim=32, n_layers=2)
    loss_fn = PhysicsLoss(m, lambda x: torch.ones_like(x), mode="fem_match")
    with torch.no_grad():
        l0 = float(loss_fn(model()))
    model.train_pde(lambda x: torch.ones_like(x), n_epochs=500, verbose=False)
    with torch.no_grad():
        assert float(loss_fn(model())) < l0
