"""CPU tests of the P2 extension (reference roadmap item, README.md:139-143 upstream): the quadrature-based oracle against
closed-form solutions and finite differences, the closed-form element matrices the CUDA kernels use (mirrored here in
Python) against the oracle's quadrature, and the host side (``FEMesh.to_p2``, the handle's patterns).  No GPU needed."""
import numpy as np
import pytest
import torch

from diffhe.mesh import FEMesh
from oracle import oracle_p2 as P

EA, EB = (0, 1, 2), (1, 2, 0)


def closed_form_tri(P3):
    """Mirror of difffe_physics_lab_b200/csrc/dfe_p2.cuh (K0 and M of one triangle)."""
    x, y = P3[:, 0], P3[:, 1]
    b = [y[1] - y[2], y[2] - y[0], y[0] - y[1]]
    c = [x[2] - x[1], x[0] - x[2], x[1] - x[0]]
    A = 0.5 * abs(c[2] * b[1] - c[1] * b[2])
    S = [[(b[i] * b[j] + c[i] * c[j]) / (4 * A) for j in range(3)] for i in range(3)]
    w = lambda p, q: 2.0 if p == q else 1.0

    def k0(I, J):
        if I < 3 and J < 3:
            return S[I][I] if I == J else -S[I][J] / 3
        if I < 3 or J < 3:
            v, e = (I, J - 3) if I < 3 else (J, I - 3)
            if EA[e] == v:
                return 4 / 3 * S[v][EB[e]]
            if EB[e] == v:
                return 4 / 3 * S[v][EA[e]]
            return 0.0
        a, bb, cc, d = EA[I - 3], EB[I - 3], EA[J - 3], EB[J - 3]
        return 4 / 3 * (S[bb][d] * w(a, cc) + S[bb][cc] * w(a, d) + S[a][d] * w(bb, cc) + S[a][cc] * w(bb, d))

    def mrow(loc):
        row = [0.0] * 6
        if loc < 3:
            for j in range(3):
                row[j] = 6.0 if j == loc else -1.0
            for k in range(3):
                row[3 + k] = -4.0 if k == (loc + 1) % 3 else 0.0
        else:
            k = loc - 3
            for j in range(3):
                row[j] = -4.0 if k == (j + 1) % 3 else 0.0
            for l in range(3):
                row[3 + l] = 32.0 if l == k else 16.0
        return row

    return np.array([[k0(i, j) for j in range(6)] for i in range(6)]), np.array([mrow(i) for i in range(6)]) * A / 180


def test_closed_form_element_matrices_match_quadrature():
    rng = np.random.default_rng(0)
    for _ in range(20):
        P3 = rng.uniform(-1, 1, (3, 2))
        nodes = np.vstack([P3, (P3[0] + P3[1]) / 2, (P3[1] + P3[2]) / 2, (P3[2] + P3[0]) / 2])
        K0, M = P.element_matrices(nodes, np.array([[0, 1, 2, 3, 4, 5]]))
        Kc, Mc = closed_form_tri(P3)
        assert np.abs(Kc - K0[0]).max() <= 1e-13 * np.abs(K0).max()
        assert np.abs(Mc - M[0]).max() <= 1e-13 * np.abs(M).max()
        assert np.abs(K0[0].sum(axis=1)).max() <= 1e-12 * np.abs(K0).max()          # constants are in the kernel
        assert abs(M[0].sum() - 0.5 * abs((P3[1, 0] - P3[0, 0]) * (P3[2, 1] - P3[0, 1]) - (P3[2, 0] - P3[0, 0]) * (P3[1, 1] - P3[0, 1]))) <= 1e-14   # sum M = area
    K0, M = P.element_matrices(np.array([[0.3], [0.9], [0.6]]), np.array([[0, 1, 2]]))
    assert np.allclose(K0[0] * 3 * 0.6, [[7, 1, -8], [1, 7, -8], [-8, -8, 16]], rtol=0, atol=1e-13)
    assert np.allclose(M[0] * 30 / 0.6, [[4, -1, 2], [-1, 4, 2], [2, 2, 16]], rtol=0, atol=1e-13)


def test_1d_quadratic_solution_is_exact():
    """-u'' = 1, u(0) = u(1) = 0: u = x(1-x)/2 lies in the P2 space (the reference's P1 test, test_fem.py:85-93 upstream)."""
    nodes, el, bc = P.line_p2(7)
    u = P.forward(nodes, el, bc, 1.0, np.ones(nodes.shape[0]))
    x = nodes[:, 0]
    assert np.abs(u - x * (1 - x) / 2).max() <= 1e-15


def test_1d_nonzero_bc_and_kappa():
    """-(kappa u')' = 0 with u(0) = 1, u(2) = 3: linear whatever the (constant) kappa."""
    nodes, el, bc = P.line_p2(5, 0.0, 2.0, 1.0, 3.0)
    u = P.forward(nodes, el, bc, 3.7, np.zeros(nodes.shape[0]))
    assert np.abs(u - (1 + nodes[:, 0])).max() <= 1e-14


def test_1d_convergence_order():
    errs = []
    for n in (8, 16, 32):
        nodes, el, bc = P.line_p2(n)
        x = nodes[:, 0]
        u = P.forward(nodes, el, bc, 1.0, np.pi ** 2 * np.sin(np.pi * x))
        errs.append(np.abs(u - np.sin(np.pi * x)).max())
    assert errs[0] / errs[1] > 7.5 and errs[1] / errs[2] > 7.5          # >= third order in the nodal max norm


def test_2d_convergence_order_and_gain_over_p1():
    errs = []
    for n in (4, 8, 16):
        m = FEMesh.rectangle(n, n).to_p2()
        x, y = m.nodes[:, 0].numpy(), m.nodes[:, 1].numpy()
        ue = np.sin(np.pi * x) * np.sin(np.pi * y)
        u = P.forward(m.nodes.numpy(), m.elements.numpy(), m.dirichlet_nodes, 1.0, 2 * np.pi ** 2 * ue)
        errs.append(np.abs(u - ue).max())
    assert errs[0] / errs[1] > 6.0 and errs[1] / errs[2] > 6.0
    assert errs[2] < 2e-4                                                # P1 on the same vertices: ~3e-3


def test_gradients_against_finite_differences():
    rng = np.random.default_rng(3)
    m = FEMesh.rectangle(3, 2, bc_value=0.3).to_p2()
    nodes, el, bc = m.nodes.numpy(), m.elements.numpy(), m.dirichlet_nodes
    kap = np.exp(rng.uniform(-1, 1, el.shape[0]))
    f = rng.uniform(-1, 1, nodes.shape[0])
    g = rng.normal(size=nodes.shape[0])
    u = P.forward(nodes, el, bc, kap, f)
    gk, gf = P.adjoint_and_grads(nodes, el, bc, kap, u, g)
    J = lambda kk, ff: float(g[m.free_nodes()] @ P.forward(nodes, el, bc, kk, ff)[m.free_nodes()])
    for e in (0, 5, 11):
        d = np.zeros_like(kap)
        d[e] = 1e-6
        assert abs((J(kap + d, f) - J(kap - d, f)) / 2e-6 - gk[e]) <= 1e-7 * np.abs(gk).sum()
    for p in (0, 7, 20, 34):
        d = np.zeros_like(f)
        d[p] = 1e-6
        assert abs((J(kap, f + d) - J(kap, f - d)) / 2e-6 - gf[p]) <= 1e-8 * np.abs(gf).max()


def test_to_p2_host_side():
    m1 = FEMesh.line(4, bc_right=None).to_p2()
    assert m1.order == 2 and m1.n_nodes == 9 and m1.elements.tolist()[1] == [1, 2, 6]
    assert m1.dirichlet_nodes == {0: 0.0}
    assert torch.equal(m1.nodes[5:, 0], 0.5 * (m1.nodes[:4, 0] + m1.nodes[1:5, 0]))
    m = FEMesh.rectangle(3, 2, bc_value=0.5).to_p2()
    assert m.n_nodes == 12 + 23 and m.n_elements == 12 and m.elements.shape[1] == 6
    assert len(m.dirichlet_nodes) == 10 + 10 and all(v == 0.5 for v in m.dirichlet_nodes.values())
    for e in m.elements.tolist():                       # m01, m12, m20 are the midpoints of their edges
        for k, (a, b) in enumerate(((0, 1), (1, 2), (2, 0))):
            assert torch.equal(m.nodes[e[3 + k]], 0.5 * (m.nodes[e[a]] + m.nodes[e[b]]))
    with pytest.raises(ValueError):
        m.to_p2()
    nm = m._native(-1)                                  # host-only handle: patterns only
    assert nm.info.n_nodes == 35 and nm.info.chain1d == 0
    rp, col = nm.csr(0)
    K, _, _, _ = P.assemble(m.nodes.numpy(), m.elements.numpy(), 1.0, np.zeros(35))
    pattern = {(int(i), int(j)) for e in m.elements.tolist() for i in e for j in e}
    assert rp[-1] == len(pattern) and all((int(p), int(c)) in pattern for p in range(35) for c in col[rp[p]:rp[p + 1]])
    assert FEMesh.rectangle(2, 2).order == 1
