"""CPU-side tests: the C ABI loads and exports what include/dfe.h declares, the symbolic (pattern)
part of the mesh handle is bit-exact against the oracle, and the host mirror of the reference API
behaves like the reference (mesh tests of upstream tests/test_fem.py:44-72 included)."""
import ctypes
import pathlib
import re

import numpy as np
import pytest
import torch

import __graft_entry__  # noqa: F401  (ensures the library is built)
from difffe_physics_lab_b200 import _native
from difffe_physics_lab_b200.solver import _kappa_mode
from diffhe.mesh import FEMesh
from diffhe.solver import DifferentiableFESolver
from oracle import oracle as O

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module", autouse=True)
def _built():
    _native.build()


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "dfe.h").read_text()
    declared = set(re.findall(r"\b(dfe_[a-z0-9_]+)\s*\(", header))
    declared -= {"dfe_status"}
    assert declared == set(_native.SYMBOLS)
    L = ctypes.CDLL(str(_native.LIB_PATH))
    for name in declared:
        assert hasattr(L, name), f"{name} not exported"
    assert _native.lib().dfe_abi_version() == 1


def test_compute_calls_fail_loudly_without_a_device():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    assert _native.lib().dfe_device_count() == 0
    m = FEMesh.line(10)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DifferentiableFESolver(m)(torch.ones(11))
    from difffe_physics_lab_b200 import assemble_sparse
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        assemble_sparse(m, 1.0, torch.ones(11))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DifferentiableFESolver(m.to_p2())(torch.ones(21))
    with pytest.raises(_native.DfeError):
        m._native(0)   # asking for a device that is not there
    # host-only handle: compute entry points refuse
    nm = m._native(-1)
    L = _native.lib()
    rc = L.dfe_solve1d_fwd(nm.handle, 1, 8, 11, 8, 0, -1, 8, 11, 8, 1 << 20, None)
    assert rc == _native.ERR_CUDA
    rc = L.dfe_assemble(nm.handle, 8, 0, 8, 8, 8, None)
    assert rc == _native.ERR_CUDA
    # batched small-system routes: not offered for a host-only handle, and the entry points refuse
    m2 = FEMesh.rectangle(5, 4)
    n2 = m2._native(-1)
    assert L.dfe_batch_supported(n2.handle) == 0 and L.dfe_band_supported(n2.handle) == 0
    assert L.dfe_batch_fwd(n2.handle, 2, 8, 30, 8, 8, 8, 8, 30, 1e-13, 100, 8, 8, 8, None) == _native.ERR_CUDA
    assert L.dfe_batch_bwd(n2.handle, 2, 8, 30, 8, 30, 8, 8, 0, 8, 30, 8, 1e-13, 100, 8, 8, 8, None) == _native.ERR_CUDA
    assert L.dfe_band_factor(n2.handle, 8, 8, None, None) == _native.ERR_CUDA
    assert L.dfe_band_fwd(n2.handle, 2, 8, 30, 8, 8, 8, 30, 8, 1 << 20, None) == _native.ERR_CUDA
    assert L.dfe_band_bwd(n2.handle, 2, 8, 30, 8, 30, 8, 0, 8, 30, 8, 8, 1 << 20, None) == _native.ERR_CUDA
    assert b"no CUDA device" in L.dfe_last_error()


# ----------------------------------------------------------------- reference mesh tests (upstream tests/test_fem.py:44-72)
def test_reference_mesh_tests():
    m = FEMesh.line(n_elements=10)
    assert (m.n_nodes, m.n_elements, m.dim) == (11, 10, 1)
    assert m.dirichlet_nodes == {0: 0.0, 10: 0.0}
    free = m.free_nodes()
    assert len(free) == 9 and 0 not in free and 10 not in free
    m = FEMesh.rectangle(nx=4, ny=4)
    assert (m.n_nodes, m.n_elements, m.dim) == (25, 32, 2)
    assert len(m.dirichlet_nodes) == 16
    assert repr(m) == "FEMesh(dim=2, n_nodes=25, n_elements=32, n_dirichlet=16)"
    assert abs(FEMesh.line(10).h() - 0.1) < 1e-15
    with pytest.raises(NotImplementedError):
        m.h()


def test_mesh_factories_bit_exact(golden):
    for name, mk in (("rect7x5_rand", lambda: FEMesh.rectangle(7, 5, (0.0, 2.3), (-1.0, 0.7), 0.3)),
                     ("rect16_ones", lambda: FEMesh.rectangle(16, 16)),
                     ("rect32_ones", lambda: FEMesh.rectangle(32, 32)),
                     ("line40_rand", lambda: FEMesh.line(40, -0.3, 2.1, -0.5, 0.25)),
                     ("line8_left_only", lambda: FEMesh.line(8, bc_left=1.0, bc_right=None)),
                     ("c1_line20", lambda: FEMesh.line(20))):
        c = golden.case(name)
        m = mk()
        assert m.nodes.dtype == torch.float64 and m.elements.dtype == torch.int64
        assert np.array_equal(m.nodes.numpy(), c["nodes"])
        assert np.array_equal(m.elements.numpy(), c["elements"])
        assert list(m.dirichlet_nodes.items()) == list(c["bc"].items())
        assert all(type(k) is int for k in m.dirichlet_nodes)


def _mesh_from_case(c):
    return FEMesh(nodes=torch.from_numpy(c["nodes"].copy()), elements=torch.from_numpy(c["elements"].copy()),
                  dirichlet_nodes=dict(c["bc"]))


@pytest.mark.parametrize("name", ["c1_line20", "line8_left_only", "line2_bc", "rect4_ones", "rect7x5_rand",
                                  "rect5x4_partial_bc", "rect3_degenerate", "rect32_ones"])
def test_csr_patterns_bit_exact_vs_oracle(golden, name):
    c = golden.case(name)
    m = _mesh_from_case(c)
    nm = m._native(-1)
    n = m.n_nodes
    rp, col = nm.csr(0)
    orp, ocol = O.structural_csr(n, c["elements"])
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    # K_free: restrict + renumber by rank in free_nodes()
    _, _, vals, F = O.assemble_csr(c["nodes"], c["elements"], 1.0, np.zeros(n))
    free, frp, fcol, _, _ = O.apply_bc(orp, ocol, vals, F, c["bc"])
    rpf, colf = nm.csr(1)
    assert np.array_equal(rpf, frp) and np.array_equal(colf, fcol)
    assert np.array_equal(nm.free_nodes(), free) and m.free_nodes() == free.tolist()
    assert nm.info.nnz_full == len(ocol) and nm.info.nnz_free == len(fcol) and nm.info.n_free == len(free)
    if "K" in c:   # pattern ⊇ nonzeros of the reference's dense K
        row_of = np.repeat(np.arange(n), np.diff(rp))
        mask = np.zeros((n, n), dtype=bool)
        mask[row_of, col] = True
        assert np.all(mask[c["K"] != 0.0])


def test_survey_nnz_counts():
    nm = FEMesh.rectangle(128, 128)._native(-1)
    assert (nm.info.n_nodes, nm.info.n_elements, nm.info.n_dirichlet, nm.info.n_free) == (16641, 32768, 512, 16129)
    assert (nm.info.nnz_full, nm.info.nnz_free, nm.info.max_row_nnz) == (115457, 111889, 7)


def test_chain_detection():
    assert FEMesh.line(20)._native(-1).info.chain1d == 1
    assert FEMesh.line(5, bc_left=None, bc_right=1.0)._native(-1).info.chain1d == 1
    assert FEMesh.line(5, bc_left=None, bc_right=None)._native(-1).info.chain1d == 0     # singular: general path reports it
    m = FEMesh.line(6)
    m.dirichlet_nodes[3] = 0.5                                                          # interior Dirichlet node
    assert m._native(-1).info.chain1d == 0
    m = FEMesh.line(6)
    m.elements = m.elements.flip(0)                                                     # element order permuted
    assert m._native(-1).info.chain1d == 0
    assert FEMesh.rectangle(3, 3)._native(-1).info.chain1d == 0


def test_native_cache_invalidation():
    m = FEMesh.line(6)
    a = m._native(-1)
    assert m._native(-1) is a
    m.dirichlet_nodes[0] = 2.0
    b = m._native(-1)
    assert b is not a
    m.nodes[3, 0] += 0.01        # in-place edit bumps the tensor version
    assert m._native(-1) is not b
    import copy
    m2 = copy.deepcopy(m)
    assert "_dfe_cache" not in m2.__dict__ and torch.equal(m2.nodes, m.nodes) and m2.dirichlet_nodes == m.dirichlet_nodes


def test_mesh_validation():
    with pytest.raises(ValueError):
        FEMesh(torch.zeros(3, 1, dtype=torch.float64), torch.tensor([[0, 5]]), {0: 0.0})._native(-1)
    with pytest.raises(ValueError):
        FEMesh(torch.zeros(3, 1, dtype=torch.float64), torch.tensor([[0, 1]]), {7: 0.0})._native(-1)
    with pytest.raises(NotImplementedError, match="Only 1D and 2D supported"):
        FEMesh(torch.zeros(4, 3, dtype=torch.float64), torch.tensor([[0, 1, 2, 3]]), {0: 0.0})._native(-1)
    with pytest.raises(NotImplementedError, match="Only 1D and 2D supported"):
        DifferentiableFESolver(FEMesh(torch.zeros(4, 3, dtype=torch.float64), torch.tensor([[0, 1, 2, 3]])))(torch.ones(4))


def test_kappa_handling_matches_reference():
    m = FEMesh.line(5)
    s = DifferentiableFESolver(m)
    assert s.kappa.dtype == torch.float64 and s.kappa.dim() == 0 and float(s.kappa) == 1.0
    assert list(s.parameters()) == []
    p64 = torch.nn.Parameter(torch.tensor(2.0, dtype=torch.float64))
    s = DifferentiableFESolver(m, kappa=p64)
    assert s.kappa is p64 and list(s.parameters()) == [p64] and "_kappa" in s.state_dict()
    p32 = torch.nn.Parameter(torch.tensor(2.0))
    s = DifferentiableFESolver(m, kappa=p32)
    assert list(s.parameters()) == [] and s.kappa.dtype == torch.float64 and s.kappa.grad_fn is not None
    k = torch.tensor(1.0, dtype=torch.float64, requires_grad=True)
    assert DifferentiableFESolver(m, kappa=k.abs()).kappa.grad_fn is not None           # demo: non-leaf kappa


def test_kappa_mode_resolution():
    K = _native
    t = torch.zeros
    assert _kappa_mode(t(()), 1, 20, False) == K.KAPPA_SCALAR
    assert _kappa_mode(t(1), 4, 20, True) == K.KAPPA_SCALAR
    assert _kappa_mode(t(20), 4, 20, True) == K.KAPPA_PER_ELEMENT
    assert _kappa_mode(t(4, 1), 4, 20, True) == K.KAPPA_PER_SAMPLE
    assert _kappa_mode(t(4, 20), 4, 20, True) == K.KAPPA_PER_SAMPLE_ELEMENT
    with pytest.raises(ValueError):
        _kappa_mode(t(5), 4, 20, True)          # the reference raises for a vector kappa too
    with pytest.raises(ValueError):
        _kappa_mode(t(4, 1), 1, 20, False)


def test_variational_loss_needs_no_gpu():
    from diffhe.loss import PhysicsLoss
    from diffhe.neural import NeuralPDE
    torch.manual_seed(0)
    m = FEMesh.line(10)
    model = NeuralPDE(m, hidden_dim=8, n_layers=2)
    u = model()
    assert abs(float(u[0])) < 1e-10 and abs(float(u[-1])) < 1e-10          # upstream test_neural.py:21-27
    loss = PhysicsLoss(m, lambda x: torch.ones_like(x), mode="variational")(u)
    assert loss.dim() == 0 and float(loss) > 0
    with pytest.raises(ValueError, match="Unknown mode"):
        PhysicsLoss(m, lambda x: x, mode="nope")
    losses = model.train_pde(lambda x: torch.ones_like(x), n_epochs=20, mode="variational", verbose=False)
    assert len(losses) == 20 and losses[-1] < losses[0]
