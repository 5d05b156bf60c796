"""The C restatement (oracle/fem_port.c, bench.py's CPU baseline) against the pinned numpy oracle."""
import numpy as np

from oracle import oracle as O
from oracle import port as P


def test_port_1d_matches_oracle(golden):
    for name in ("c1_line20", "line40_rand", "line8_left_only", "line8_right_only", "line10_bc12_rand", "line3_bc"):
        c = golden.case(name)
        u, gk, gf = P.solve1d_batch(c["nodes"][:, 0], c["bc"], c["f"][None], float(c["kappa"]), c["gbar"][None])
        assert np.abs(u[0] - c["u"]).max() <= 1e-12 * max(np.abs(c["u"]).max(), 1e-300)
        go, gfo, _ = O.adjoint_and_grads(c["nodes"], c["elements"], c["bc"], float(c["kappa"]), c["u"], c["gbar"])
        assert abs(gk[0] - go.sum()) <= 1e-11 * np.abs(go).sum()
        assert np.abs(gf[0] - gfo).max() <= 1e-11 * max(np.abs(gfo).max(), 1e-300)


def test_port_1d_batch_threads():
    rng = np.random.default_rng(0)
    nodes, el, bc = O.line_mesh(3000)
    f = rng.uniform(0, 1, (6, 3001))
    kap = rng.uniform(0.5, 2, 6)
    gbar = rng.standard_normal((6, 3001))
    u, gk, gf = P.solve1d_batch(nodes[:, 0], bc, f, kap, gbar)
    u1, gk1, gf1 = P.solve1d_batch(nodes[:, 0], bc, f, kap, gbar, nthreads=1)
    assert np.array_equal(u, u1) and np.array_equal(gk, gk1) and np.array_equal(gf, gf1)
    uo = O.forward(nodes, el, bc, float(kap[2]), f[2], exact=False)
    assert np.abs(u[2] - uo).max() <= 1e-9 * np.abs(uo).max()     # plain float64 Thomas vs banded LU: both ~1e-11 here


def test_port_pcg_matches_numpy_pcg():
    nodes, el, bc = O.rectangle_mesh(24, 24)
    rowptr, col, vals, F = O.assemble_csr(nodes, el, 1.0, np.ones(625))
    _, rp, cf, vf, Ff = O.apply_bc(rowptr, col, vals, F, bc)
    x, it, rel = P.pcg_csr(rp, cf, vf, Ff)
    xo, ito, _ = O.jacobi_pcg(rp, cf, vf, Ff)
    assert abs(it - ito) <= 2 and rel <= 1e-13
    assert np.abs(x - xo).max() <= 1e-12 * np.abs(xo).max()
