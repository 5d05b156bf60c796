"""GPU parity tests of the pipelined 1-D kernel IN THE REGIME THE BENCH RUNS IT (run with ``-m gpu``).

``k1d_pipe`` walks a batch in iterations: group g of CTAs handles samples g, g+NG, g+2NG, ...; ring slots are reused
every NR = LB+LC+3 iterations, the mbarrier phases wrap, and every iteration works on three different samples
(phases A / B / C lag by LB and LC).  The tests below force many iterations per group (B >> NG) on single-chunk and
multi-chunk meshes, odd and even row pitches, every boundary-condition type, per-sample and shared kappa (the ``SK``
adjoint specialisation), forward and adjoint, and compare

* rows taken from every pipeline position (iterations 0, 1, LB, LB+LC, NR-1, NR, NR+1, 2NR, last) against the exact
  oracle (the same float64 system solved in 50-digit arithmetic) at 1e-12, and
* ALL rows against the two-pass split kernels (no cross-CTA pipeline at all) at 1e-13.

Also here: the fused misfit adjoint (config 5 step), the host <-> device row streaming of the product, the multi-sweep
kernel (n_refine != 1, meshes above 2e5 nodes) and the size limit of the fused path.
"""
import math

import numpy as np
import pytest
import torch

from difffe_physics_lab_b200 import _native
from diffhe.loss import PhysicsLoss
from diffhe.mesh import FEMesh
from diffhe.solver import DifferentiableFESolver
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL1D = 1e-12
LB, LC, NR = 2, 2, 7                      # default configuration of k1d_pipe (dfe_1d_pipe.cu)
CAP_FWD, CAP_BWD = 9 * 32 * 12, 11 * 32 * 8


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _native.build()


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def groups_resident(nn, B, cap):
    """NG of run_cfg (dfe_1d_pipe.cu): one CTA per SM, gmin chunks per sample."""
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    gmin = -(-nn // cap)
    return max(1, min(B, sm // gmin))


def pipeline_positions(B, NG):
    """One sample from each interesting iteration of the pipeline (different groups), plus the last sample."""
    n_it = -(-B // NG)
    its = sorted({0, 1, LB, LB + LC, NR - 1, NR, NR + 1, 2 * NR, n_it // 2, n_it - 2, n_it - 1} & set(range(n_it)))
    out = []
    for k, it in enumerate(its):
        s = it * NG + (3 * k + 1) % NG
        if s < B:
            out.append(s)
    return sorted(set(out + [B - 1]))


def abi_split(mesh, f, kap, mode, gbar):
    """Forward + adjoint through the C ABI on rows that start on an odd 8-byte offset: the pipelined kernel refuses
    those (aligned-superset TMA loads), so this runs the two-pass split kernels on the same data."""
    L = _native.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    nm = mesh._native(dev.index)
    B, n = f.shape

    def shifted(t):
        big = torch.empty(t.numel() + 1, dtype=torch.float64, device=dev)
        v = big[1:].view(t.shape)
        v.copy_(t)
        assert v.data_ptr() % 16 == 8
        return v

    fs, gs = shifted(f), shifted(gbar)
    u, gf = shifted(torch.zeros_like(f)), shifted(torch.zeros_like(f))
    gk = torch.zeros_like(kap)
    ws = torch.empty(L.dfe_solve1d_workspace_bytes(nm.handle, B), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _native.check(L.dfe_solve1d_fwd(nm.handle, B, fs.data_ptr(), n, kap.data_ptr(), mode, 1, u.data_ptr(), n, ws.data_ptr(), ws.numel(), st))
    _native.check(L.dfe_solve1d_bwd(nm.handle, B, gs.data_ptr(), n, u.data_ptr(), n, kap.data_ptr(), mode, 1, gf.data_ptr(), n,
                                    gk.data_ptr(), ws.data_ptr(), ws.numel(), st))
    torch.cuda.synchronize()
    return u, gf, gk


CASES = [
    # n_el, B, (bc_left, bc_right), kappa layout
    (2000, 600, (0.0, 0.0), "per_sample"),       # 1 chunk per sample, NG = #SMs: ~5 iterations, odd row pitch (2001)
    (2000, 1500, (0.3, None), "shared"),         # natural right end, shared kappa (SK adjoint), ~11 iterations
    (2001, 1200, (None, -0.2), "per_sample"),    # natural left end, even row pitch (2002)
    (16384, 700, (0.0, 0.0), "shared"),          # config 5a geometry (5-7 chunks per sample), ~30 iterations
    (16385, 400, (0.25, -0.5), "per_sample"),    # even pitch, non-zero Dirichlet data on both ends
    (100000, 64, (0.1, -0.4), "per_sample"),     # config 2 geometry (29 / 36 chunks per sample), 13 / 16 iterations
    (100000, 48, (0.0, 0.0), "shared"),
]


@pytest.mark.parametrize("n,B,bcs,klay", CASES)
def test_pipe_many_iterations_vs_oracle_and_split(n, B, bcs, klay):
    rng = np.random.default_rng(1000 + n + B)
    dev = "cuda"
    m = FEMesh.line(n, x_left=-0.3, x_right=1.1, bc_left=bcs[0], bc_right=bcs[1])
    nn = n + 1
    f = torch.tensor(rng.uniform(0, 1, (B, nn)), device=dev)
    gbar = torch.tensor(rng.standard_normal((B, nn)), device=dev)
    if klay == "per_sample":
        kap = torch.tensor(np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1))), device=dev)
        mode = _native.KAPPA_PER_SAMPLE
    else:
        kap = torch.tensor(1.37, dtype=torch.float64, device=dev)
        mode = _native.KAPPA_SCALAR
    kr = kap.clone().requires_grad_(True)
    fr = f.clone().requires_grad_(True)
    u = DifferentiableFESolver(m, kappa=kr)(fr)
    u.backward(gbar)
    torch.cuda.synchronize()
    assert _native.lib().dfe_mesh_fault(m._native(torch.cuda.current_device()).handle) == 0
    assert groups_resident(nn, B, CAP_FWD) < B          # really more than one iteration per group
    # ---- every row against the split kernels
    us, gfs, gks = abi_split(m, f, kap.reshape(-1).contiguous(), mode, gbar)
    assert float((u.detach() - us).abs().max()) <= 1e-13 * float(us.abs().max())
    assert float((fr.grad - gfs).abs().max()) <= 1e-13 * float(gfs.abs().max())
    # ---- rows from every pipeline position against the exact oracle
    rows = sorted(set(pipeline_positions(B, groups_resident(nn, B, CAP_FWD)) + pipeline_positions(B, groups_resident(nn, B, CAP_BWD))))
    if n >= 100000:
        rows = rows[::2] + [rows[-1]]
    nodes, el, bc = m.nodes.numpy(), m.elements.numpy(), m.dirichlet_nodes
    un, gfn = u.detach().cpu().numpy(), fr.grad.cpu().numpy()
    fn, gn, kn = f.cpu().numpy(), gbar.cpu().numpy(), kap.detach().cpu().numpy().reshape(-1)
    mag = 0.0
    for b in rows:
        kb = float(kn[b] if klay == "per_sample" else kn[0])
        uo = O.forward(nodes, el, bc, kb, fn[b])
        gko, gfo, _ = O.adjoint_and_grads(nodes, el, bc, kb, uo, gn[b])
        assert relerr(un[b], uo) <= TOL1D, f"u, sample {b}"
        assert np.abs(gfn[b] - gfo).max() <= TOL1D * np.abs(gfo).max(), f"dL/df, sample {b}"
        mag += np.abs(gko).sum()
        if klay == "per_sample":
            assert abs(float(kr.grad[b, 0]) - gko.sum()) <= TOL1D * np.abs(gko).sum(), f"dL/dkappa, sample {b}"
    mag *= B / len(rows)                                  # estimate of sum_b sum_e |dL/dkappa_e|
    if klay == "per_sample":
        assert float((kr.grad.reshape(-1) - gks).abs().max()) <= 1e-11 * mag / B
    else:
        assert abs(float(kr.grad) - float(gks)) <= TOL1D * mag
    # ---- adjoint without dL/df (f is data): the kernel variant that writes no row and frees its ring slots after phase B
    kn2 = kap.clone().requires_grad_(True)
    DifferentiableFESolver(m, kappa=kn2)(f).backward(gbar)
    torch.cuda.synchronize()
    assert _native.lib().dfe_mesh_fault(m._native(torch.cuda.current_device()).handle) == 0
    if klay == "per_sample":
        for b in rows:
            gko, _, _ = O.adjoint_and_grads(nodes, el, bc, float(kn[b]), O.forward(nodes, el, bc, float(kn[b]), fn[b]), gn[b])
            assert abs(float(kn2.grad[b, 0]) - gko.sum()) <= TOL1D * np.abs(gko).sum(), f"dL/dkappa (no dL/df), sample {b}"
        assert float((kn2.grad - kr.grad).abs().max()) <= 1e-11 * mag / B
    else:
        assert abs(float(kn2.grad) - float(kr.grad)) <= TOL1D * mag


@pytest.mark.parametrize("n,B,klay", [(2000, 600, "shared"), (2001, 500, "per_sample"), (16384, 300, "shared")])
def test_fused_misfit_adjoint(n, B, klay):
    """solver.misfit(f, u_data) (forward solve + dfe_solve1d_bwd_misfit) against the composed torch expression and the
    exact oracle: the loss of the reference's kappa-recovery loop (examples/poisson_1d_demo.py:104-110), batched."""
    rng = np.random.default_rng(77 + n)
    dev = "cuda"
    m = FEMesh.line(n, bc_left=0.1, bc_right=-0.2)
    nn = n + 1
    f = torch.tensor(rng.uniform(0.5, 1.5, (B, nn)), device=dev)
    with torch.no_grad():
        u_data = DifferentiableFESolver(m, kappa=torch.tensor(2.0, dtype=torch.float64, device=dev))(f)
        u_data = u_data + 1e-3 * torch.tensor(rng.standard_normal((B, nn)), device=dev)     # noisy data, BC nodes too
    if klay == "shared":
        k0 = torch.tensor(1.0, dtype=torch.float64, device=dev)
    else:
        k0 = torch.tensor(np.exp(rng.uniform(np.log(0.6), np.log(1.8), (B, 1))), device=dev)
    # fused
    k1 = k0.clone().requires_grad_(True)
    f1 = f.clone().requires_grad_(True)
    s1 = DifferentiableFESolver(m, kappa=k1)
    loss1 = s1.misfit(f1, u_data)
    loss1.backward()
    # composed: ordinary forward, torch loss, ordinary adjoint
    k2 = k0.clone().requires_grad_(True)
    f2 = f.clone().requires_grad_(True)
    u2 = DifferentiableFESolver(m, kappa=k2)(f2)
    loss2 = ((u2 - u_data) ** 2).sum() / nn
    loss2.backward()
    assert torch.equal(s1._opts["last_u"], u2.detach())                     # same forward kernel, same bits
    assert abs(float(loss1) - float(loss2)) <= 1e-13 * abs(float(loss2))
    gmag = float(k2.grad.abs().sum())
    assert float((k1.grad - k2.grad).abs().max()) <= 1e-10 * gmag / max(1, k2.grad.numel())
    assert float((f1.grad - f2.grad).abs().max()) <= 1e-12 * float(f2.grad.abs().max())
    # oracle on a few rows
    nodes, el, bc = m.nodes.numpy(), m.elements.numpy(), m.dirichlet_nodes
    NG = groups_resident(nn, B, CAP_BWD)
    rows = pipeline_positions(B, NG)[:6]
    kn = k0.cpu().numpy().reshape(-1)
    tot = 0.0
    for b in rows:
        kb = float(kn[b] if klay == "per_sample" else kn[0])
        uo = O.forward(nodes, el, bc, kb, f[b].cpu().numpy())
        d = u_data[b].cpu().numpy()
        gko, gfo, _ = O.adjoint_and_grads(nodes, el, bc, kb, uo, 2.0 * (uo - d) / nn)
        assert np.abs(f1.grad[b].cpu().numpy() - gfo).max() <= 1e-11 * np.abs(gfo).max()
        if klay == "per_sample":
            assert abs(float(k1.grad[b, 0]) - gko.sum()) <= 1e-11 * np.abs(gko).sum()
        tot += ((uo - d) ** 2).sum() / nn
    # without dL/df (the inverse-problem step): no row is written and the kernel runs with a different chunk geometry
    # (ring slots are released after phase B), so the sums are the same numbers in a different fixed order
    k3 = k0.clone().requires_grad_(True)
    loss3 = DifferentiableFESolver(m, kappa=k3).misfit(f, u_data)
    loss3.backward()
    assert abs(float(loss3) - float(loss1)) <= 1e-13 * abs(float(loss1))
    assert float((k3.grad - k1.grad).abs().max()) <= 1e-10 * gmag / max(1, k2.grad.numel())
    k3b = k0.clone().requires_grad_(True)
    loss3b = DifferentiableFESolver(m, kappa=k3b).misfit(f, u_data)
    loss3b.backward()
    assert torch.equal(loss3b.detach(), loss3.detach()) and torch.equal(k3b.grad, k3.grad)     # bit-reproducible
    # shared kappa: the two words land in a caller-provided buffer (the all-reduce buffer of the sharded step)
    if klay == "shared":
        out2 = torch.zeros(2, dtype=torch.float64, device=dev)
        k4 = k0.clone().requires_grad_(True)
        loss4 = DifferentiableFESolver(m, kappa=k4).misfit(f, u_data, out2=out2)
        assert float(out2[1]) == float(loss4) == float(loss3) and float(out2[0]) == float(k3.grad)


def test_misfit_demo_step0(golden):
    """Demo step 0 (examples/poisson_1d_demo.py:104-110): loss 0.002016126543209874, grad -0.008064506172839506."""
    d = golden.case("demo_line30")
    m = FEMesh.line(30)
    f = torch.ones(31, dtype=torch.float64, device="cuda")
    with torch.no_grad():
        u_data = DifferentiableFESolver(m, kappa=torch.tensor(2.0, dtype=torch.float64))(f)
    k = torch.tensor(1.0, dtype=torch.float64, requires_grad=True)
    loss = DifferentiableFESolver(m, kappa=k.abs()).misfit(f, u_data)
    loss.backward()
    assert abs(float(loss) - float(d["loss"])) < 1e-16
    assert abs(float(k.grad) - float(d["gkappa"])) < 1e-15


@pytest.mark.parametrize("pinned", [False, True])
def test_host_rows_are_streamed(pinned):
    """A large host-resident batch: f (and gbar) stream host -> device in row chunks while earlier chunks are solved,
    u and dL/df stream back; results equal the device-resident run to rounding (chunking changes NG / the chunk count
    per sample, i.e. summation orders), CPU in -> CPU out."""
    rng = np.random.default_rng(21)
    n, B = 20000, 320                                       # 51 MB per array: 3 chunks
    m = FEMesh.line(n, bc_left=0.2, bc_right=None)
    f = torch.tensor(rng.uniform(0, 1, (B, n + 1)))
    gbar = torch.tensor(rng.standard_normal((B, n + 1)))
    kap = torch.tensor(np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1))))
    if pinned:
        f, gbar = f.pin_memory(), gbar.pin_memory()
    # device-resident reference run
    kd = kap.cuda().requires_grad_(True)
    fd = f.cuda().requires_grad_(True)
    ud = DifferentiableFESolver(m, kappa=kd)(fd)
    ud.backward(gbar.cuda())
    # host run: CPU in, CPU out, CPU gradients
    kh = kap.clone().requires_grad_(True)
    fh = f.clone().requires_grad_(True) if not pinned else f.requires_grad_(True)
    uh = DifferentiableFESolver(m, kappa=kh)(fh)
    assert uh.device.type == "cpu" and uh.shape == (B, n + 1)
    uh.backward(gbar)
    assert fh.grad.device.type == "cpu" and kh.grad.device.type == "cpu"
    assert float((uh.detach() - ud.detach().cpu()).abs().max()) <= 1e-13 * float(ud.abs().max())
    assert float((fh.grad - fd.grad.cpu()).abs().max()) <= 1e-13 * float(fd.grad.abs().max())
    assert float((kh.grad - kd.grad.cpu()).abs().max()) <= 1e-10 * float(kd.grad.abs().max())
    # host f, solution kept on the device (out_device), shared kappa: dL/dkappa is the chunk-ordered sum
    ks = torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
    us = DifferentiableFESolver(m, kappa=ks, out_device="cuda")(f.detach())
    assert us.is_cuda
    us.backward(gbar.cuda())
    ks2 = torch.tensor(1.3, dtype=torch.float64, device="cuda", requires_grad=True)
    us2 = DifferentiableFESolver(m, kappa=ks2)(f.detach().cuda())
    us2.backward(gbar.cuda())
    assert float((us - us2).abs().max()) <= 1e-13 * float(us2.abs().max())
    assert abs(float(ks.grad) - float(ks2.grad)) <= 1e-9 * abs(float(ks2.grad))
    us3 = DifferentiableFESolver(m, kappa=1.3, out_device="cuda")(f.detach())
    assert torch.equal(us3, us)                               # deterministic


@pytest.mark.parametrize("n_refine,tol", [(0, 1e-6), (2, TOL1D), (3, TOL1D)])
def test_multi_sweep_kernel_n_refine(n_refine, tol):
    """The multi-sweep kernel (k_solve1d, used for n_refine != 1): forward and adjoint against the exact oracle.
    n_refine = 0 is the structured solve alone (error ~ |M^-1 E| ~ n^2 eps), 2 and 3 are exact to rounding."""
    rng = np.random.default_rng(31 + n_refine)
    n, B = 9000, 40                                           # 3 chunks per sample, more samples than groups? (NG ~ 98)
    m = FEMesh.line(n, bc_left=0.3, bc_right=-0.1)
    f = rng.uniform(0, 1, (B, n + 1))
    gbar = rng.standard_normal((B, n + 1))
    kap = np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1)))
    k = torch.tensor(kap, device="cuda", requires_grad=True)
    ft = torch.tensor(f, device="cuda", requires_grad=True)
    u = DifferentiableFESolver(m, kappa=k, n_refine=n_refine)(ft)
    u.backward(torch.tensor(gbar, device="cuda"))
    nodes, el, bc = m.nodes.numpy(), m.elements.numpy(), m.dirichlet_nodes
    for b in (0, 7, B - 1):
        uo = O.forward(nodes, el, bc, float(kap[b, 0]), f[b])
        gko, gfo, _ = O.adjoint_and_grads(nodes, el, bc, float(kap[b, 0]), uo, gbar[b])
        assert relerr(u[b].detach().cpu().numpy(), uo) <= tol
        assert np.abs(ft.grad[b].cpu().numpy() - gfo).max() <= tol * np.abs(gfo).max()
        assert abs(float(k.grad[b, 0]) - gko.sum()) <= max(tol, 1e-11) * np.abs(gko).sum()


def test_large_chain_two_sweeps_and_size_limit():
    """Above 2e5 nodes the fused path runs two Neumann sweeps (multi-sweep kernel, cooperative launch); beyond the
    co-resident capacity of that kernel the solver refuses the mesh in forward (decided once, for both directions)."""
    rng = np.random.default_rng(41)
    n, B = 250000, 150                                        # more samples than resident groups
    m = FEMesh.line(n)
    f = rng.uniform(0, 1, (B, n + 1))
    gbar = rng.standard_normal((B, n + 1))
    k = torch.tensor(np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1))), device="cuda", requires_grad=True)
    ft = torch.tensor(f, device="cuda", requires_grad=True)
    u = DifferentiableFESolver(m, kappa=k)(ft)
    u.backward(torch.tensor(gbar, device="cuda"))
    nodes, el, bc = m.nodes.numpy(), m.elements.numpy(), m.dirichlet_nodes
    for b in (3, B - 1):
        kb = float(k[b, 0])
        uo = O.forward(nodes, el, bc, kb, f[b])
        gko, gfo, _ = O.adjoint_and_grads(nodes, el, bc, kb, uo, gbar[b])
        assert relerr(u[b].detach().cpu().numpy(), uo) <= TOL1D
        assert np.abs(ft.grad[b].cpu().numpy() - gfo).max() <= TOL1D * np.abs(gfo).max()
        assert abs(float(k.grad[b, 0]) - gko.sum()) <= 1e-11 * np.abs(gko).sum()
    L = _native.lib()
    dev = torch.cuda.current_device()
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    cap = 2 * sm * 13 * 256                                   # chunks of the adjoint kernel that fit at once
    ok = FEMesh.line(cap - 1)
    big = FEMesh.line(cap)
    assert L.dfe_solve1d_supported(ok._native(dev).handle, _native.KAPPA_SCALAR, -1) == 1
    assert L.dfe_solve1d_supported(big._native(dev).handle, _native.KAPPA_SCALAR, -1) == 0
    assert L.dfe_solve1d_supported(big._native(dev).handle, _native.KAPPA_SCALAR, 1) == 1   # one sweep: split kernels, any size
    with pytest.raises(NotImplementedError):
        DifferentiableFESolver(big)(torch.ones(cap + 1, dtype=torch.float64, device="cuda"))
    # at the limit both directions run
    kk = torch.tensor(1.0, dtype=torch.float64, device="cuda", requires_grad=True)
    uu = DifferentiableFESolver(ok, kappa=kk)(torch.ones(cap, dtype=torch.float64, device="cuda"))
    uu.sum().backward()
    # (the float64 system of a ~1e6-node chain is ~3e-6 away from the analytic x(1-x)/2: its own rounding, |M^-1 E| ~ 1e-5;
    # the kernel has to reproduce THAT system's solution, which only the exact oracle knows)
    uo = O.forward(ok.nodes.numpy(), ok.elements.numpy(), ok.dirichlet_nodes, 1.0, np.ones(cap))
    assert relerr(uu.detach().cpu().numpy(), uo) <= TOL1D
    assert abs(float(kk.grad) + float(uu.sum())) <= 1e-4 * abs(float(uu.sum()))


def test_physics_loss_memo_hit_and_invalidation(monkeypatch):
    """PhysicsLoss memoises the FEM target: same forcing, kappa and mesh -> no second solve; an in-place
    kappa update, a different forcing or a mesh edit -> a new solve."""
    m = FEMesh.line(16)
    kap = torch.tensor(1.5, dtype=torch.float64)
    solver = DifferentiableFESolver(m, kappa=kap)
    calls = []
    orig = DifferentiableFESolver.forward

    def counting(self, f):
        calls.append(1)
        return orig(self, f)

    monkeypatch.setattr(DifferentiableFESolver, "forward", counting)
    amp = [1.0]
    loss = PhysicsLoss(m, lambda x: amp[0] * torch.ones_like(x), mode="fem_match", solver=solver)
    up = torch.zeros(17, dtype=torch.float64)
    l0 = float(loss(up))
    l1 = float(loss(up))
    assert len(calls) == 1 and l0 == l1                      # hit
    with torch.no_grad():
        kap.mul_(2.0)                                        # in-place update (what an optimiser step does)
    l2 = float(loss(up))
    assert len(calls) == 2 and abs(l2 - l0 / 4.0) <= 1e-12 * l0   # u ~ 1/kappa
    amp[0] = 2.0
    l3 = float(loss(up))
    assert len(calls) == 3 and abs(l3 - l0) <= 1e-12 * l0
    m.dirichlet_nodes[16] = 0.5                              # mesh edit -> fingerprint changes
    loss(up)
    assert len(calls) == 4
    m.dirichlet_nodes[16] = -1.0
    loss(up)
    m.dirichlet_nodes[16] = -2.0                             # hash(-1.0) == hash(-2.0) in CPython: must still miss
    loss(up)
    assert len(calls) == 6
