"""Pin the CPU oracle (oracle/oracle.py) against the unmodified reference.

Golden vectors come from tests/golden/make_golden.py (reference run in the build
container).  Dense K/F must match BIT-EXACTLY; solutions and gradients within the
reference's own LU round-off.
"""
import numpy as np
import pytest

from oracle import oracle as O

DENSE_CASES = ["c1_line20", "line10_bc12_rand", "line40_rand", "line8_left_only", "line8_right_only",
               "line2_bc", "line3_bc", "rect4_ones", "rect8_ones", "rect7x5_rand", "rect6x9_rand",
               "rect5x4_partial_bc", "rect3_degenerate"]
GRAD_CASES = DENSE_CASES + ["line100_ones", "line400_ones", "rect16_ones"]


@pytest.mark.parametrize("name", DENSE_CASES)
def test_dense_K_F_bit_exact(golden, name):
    c = golden.case(name)
    K, F = O.assemble_dense(c["nodes"], c["elements"], float(c["kappa"]), c["f"])
    assert np.array_equal(K, c["K"]), "dense K differs from the reference bitwise"
    assert np.array_equal(F, c["F"]), "F differs from the reference bitwise"


@pytest.mark.parametrize("name", DENSE_CASES)
def test_csr_pattern_and_values(golden, name):
    c = golden.case(name)
    n = c["nodes"].shape[0]
    rowptr, col, vals, F = O.assemble_csr(c["nodes"], c["elements"], float(c["kappa"]), c["f"])
    row_of = np.repeat(np.arange(n), np.diff(rowptr))
    # rows ascending, columns strictly ascending within a row
    for r in range(n):
        cc = col[rowptr[r]:rowptr[r + 1]]
        assert np.all(np.diff(cc) > 0)
    dense = np.zeros((n, n))
    dense[row_of, col] = vals
    assert np.array_equal(dense, c["K"])            # values bit-exact at pattern, exact zeros elsewhere
    assert np.array_equal(F, c["F"])
    mask = np.zeros((n, n), dtype=bool)
    mask[row_of, col] = True
    assert np.all(mask[c["K"] != 0.0])              # pattern ⊇ nonzeros of the reference K


@pytest.mark.parametrize("name", GRAD_CASES)
def test_forward_and_grads(golden, name):
    c = golden.case(name)
    u = O.forward(c["nodes"], c["elements"], c["bc"], float(c["kappa"]), c["f"])
    scale = max(np.abs(c["u"]).max(), 1e-300)
    assert np.abs(u - c["u"]).max() <= 2e-12 * scale
    gk, gf, _ = O.adjoint_and_grads(c["nodes"], c["elements"], c["bc"], float(c["kappa"]), u, c["gbar"])
    # tolerance relative to sum_e |dL/dkappa_e|: the sum cancels
    assert abs(gk.sum() - float(c["gkappa"])) <= 1e-10 * max(np.abs(gk).sum(), 1e-300)
    assert np.abs(gf - c["gf"]).max() <= 1e-11 * max(np.abs(c["gf"]).max(), 1e-300)


def test_reference_known_answers(golden):
    """The pins of the reference's own tests (tests/test_fem.py:85-179 upstream)."""
    for n, atol in ((10, 1e-10), (100, 1e-9)):
        nodes, el, bc = O.line_mesh(n)
        u = O.forward(nodes, el, bc, 1.0, np.ones(n + 1))
        x = nodes[:, 0]
        assert np.allclose(u, x * (1 - x) / 2, atol=atol)
        assert abs(u[0]) < 1e-12 and abs(u[-1]) < 1e-12
    c = golden.case("line10_bc12_f0")
    u = O.forward(c["nodes"], c["elements"], c["bc"], 1.0, c["f"])
    assert np.allclose(u, 1.0 + c["nodes"][:, 0], atol=1e-10)
    assert np.abs(u - c["u"]).max() < 1e-13
    errs = []
    for n in (10, 20, 40, 80):
        nodes, el, bc = O.line_mesh(n)
        x = nodes[:, 0]
        u = O.forward(nodes, el, bc, 1.0, np.pi ** 2 * np.sin(np.pi * x))
        errs.append(np.abs(u - np.sin(np.pi * x)).max())
    assert all(errs[i - 1] / errs[i] > 3.0 for i in range(1, 4))
    nodes, el, bc = O.rectangle_mesh(4, 4)
    assert np.abs(O.forward(nodes, el, bc, 1.0, np.zeros(25))).max() < 1e-10
    nodes, el, bc = O.rectangle_mesh(8, 8)
    u = O.forward(nodes, el, bc, 1.0, np.ones(81))
    assert u[O.free_nodes(81, bc)].min() > 0


def test_survey_constants(golden):
    """SURVEY §8(c) constants (captured from the unmodified reference)."""
    c = golden.case("c1_line20")
    assert float(c["gkappa"]) == float.fromhex("-0x1.a99999999998ep+0")
    assert float(c["u"][10]) == float.fromhex("0x1.ffffffffffff9p-4")
    d = golden.case("demo_line30")
    assert float(d["loss"]) == 0.002016126543209874
    assert float(d["gkappa"]) == -0.008064506172839506
    u_data = O.forward(d["nodes"], d["elements"], d["bc"], 2.0, np.ones(31))
    u = O.forward(d["nodes"], d["elements"], d["bc"], 1.0, np.ones(31))
    loss = ((u - u_data) ** 2).mean()
    assert abs(loss - float(d["loss"])) < 1e-15
    gk, _, _ = O.adjoint_and_grads(d["nodes"], d["elements"], d["bc"], 1.0, u, 2 * (u - u_data) / 31)
    assert abs(gk.sum() - float(d["gkappa"])) < 1e-14
    for n, mx in ((4, 0.070312499999999986), (8, 0.072782628676470576), (16, 0.073445766578919713)):
        assert abs(golden.case(f"rect{n}_ones")["u"].max() - mx) < 1e-15


def test_mesh_generators_bit_exact(golden):
    for name, args in (("rect7x5_rand", (7, 5, (0.0, 2.3), (-1.0, 0.7), 0.3)), ("rect16_ones", (16, 16)), ("rect32_ones", (32, 32))):
        c = golden.case(name)
        nodes, el, bc = O.rectangle_mesh(*args)
        assert np.array_equal(nodes, c["nodes"]) and np.array_equal(el, c["elements"])
        assert list(bc.keys()) == list(c["bc"].keys()) and list(bc.values()) == list(c["bc"].values())
    c = golden.case("line40_rand")
    nodes, el, bc = O.line_mesh(40, -0.3, 2.1, -0.5, 0.25)
    assert np.array_equal(nodes, c["nodes"]) and np.array_equal(el, c["elements"]) and bc == c["bc"]


def test_forward_32_and_f32(golden):
    c = golden.case("rect32_ones")
    u = O.forward(c["nodes"], c["elements"], c["bc"], 1.0, c["f"])
    assert np.abs(u - c["u"]).max() <= 1e-12 * c["u"].max()
    c = golden.case("line12_f32")
    u = O.forward(c["nodes"], c["elements"], c["bc"], 1.0, c["f"].astype(np.float64))
    assert np.abs(u - c["u"]).max() <= 1e-14
    c = golden.case("line1600_ones")
    u = O.forward(c["nodes"], c["elements"], c["bc"], 1.0, c["f"])
    assert np.abs(u - c["u"]).max() <= 5e-12 * c["u"].max()   # reference's own dense-LU error ~1e-12 here


def test_per_element_kappa_extension():
    """rectangle(4,4), kappa_e = 1 + 0.5 e/32, f=1, J = F^T u."""
    nodes, el, bc = O.rectangle_mesh(4, 4)
    kap = 1.0 + 0.5 * np.arange(32) / 32
    f = np.ones(25)
    u = O.forward(nodes, el, bc, kap, f)
    _, _, _, F = O.assemble_csr(nodes, el, kap, f)
    J = F @ u
    assert abs(J - 0.02336457495061802) < 1e-16
    gbar = F.copy()
    gk, _, _ = O.adjoint_and_grads(nodes, el, bc, kap, u, gbar)
    ref8 = [0, -0.00147538203336745, -0.00073769101668373, -0.00118563863606154,
            -0.00114170043120912, -0.00073796675704933, -0.00067784987673196, -0.00067784987673196]
    assert np.abs(gk[:8] - ref8).max() < 1e-15
    assert abs(gk.sum() - (-0.019090990124307906)) < 1e-15


def test_jacobi_pcg_matches_direct():
    nodes, el, bc = O.rectangle_mesh(32, 32)
    rowptr, col, vals, F = O.assemble_csr(nodes, el, 1.0, np.ones(33 * 33))
    free, rp, cf, vf, Ff = O.apply_bc(rowptr, col, vals, F, bc)
    x, it, rel = O.jacobi_pcg(rp, cf, vf, Ff)
    xd = O.solve_csr(rp, cf, vf, Ff)
    assert it == 72                     # SURVEY §7 hard part 2
    assert np.abs(x - xd).max() <= 1e-12 * np.abs(xd).max()
