"""ctypes wrapper of oracle/fem_port.c (the timed CPU baseline).  TEST / MEASUREMENT INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = HERE / "libfem_port.so"
        if not so.exists() or so.stat().st_mtime < (HERE / "fem_port.c").stat().st_mtime:
            subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
        L = C.CDLL(str(so))
        L.port_max_threads.restype = C.c_int
        L.port_solve1d_batch.restype = C.c_int
        L.port_solve1d_batch.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_long,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int]
        L.port_pcg_csr.restype = C.c_long
        L.port_pcg_csr.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                   C.c_long, C.POINTER(C.c_double), C.c_int]
        L.port_pcg_csr_batch.restype = C.c_long
        L.port_pcg_csr_batch.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
                                         C.c_double, C.c_long, C.c_int]
        _lib = L
    return _lib


def max_threads() -> int:
    return lib().port_max_threads()


def solve1d_batch(x, bc, f, kappa, gbar=None, need_gf=True, nthreads=0):
    """Batched 1-D forward (+ adjoint).  x (nn,), f (B, nn), kappa scalar or (B,), gbar (B, nn) or None."""
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    f = np.ascontiguousarray(f, dtype=np.float64)
    B, nn = f.shape
    kap = np.ascontiguousarray(np.atleast_1d(np.asarray(kappa, dtype=np.float64)).reshape(-1))
    per_sample = int(kap.size == B and B > 1)
    u = np.empty_like(f)
    bcL, bcR = 0 in bc, (nn - 1) in bc
    gL, gR = float(bc.get(0, 0.0)), float(bc.get(nn - 1, 0.0))
    gk = gf = None
    if gbar is not None:
        gbar = np.ascontiguousarray(gbar, dtype=np.float64)
        gk = np.empty(B)
        gf = np.empty_like(f) if need_gf else None
    rc = lib().port_solve1d_batch(nn, x.ctypes.data, int(bcL), gL, int(bcR), gR, B, f.ctypes.data, kap.ctypes.data,
                                  per_sample, gbar.ctypes.data if gbar is not None else None, u.ctypes.data,
                                  gf.ctypes.data if gf is not None else None, gk.ctypes.data if gk is not None else None,
                                  nthreads)
    if rc:
        raise RuntimeError("port_solve1d_batch failed")
    return u, gk, gf


def pcg_csr(rowptr, col, vals, b, tol=1e-13, maxit=100000, nthreads=0):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty_like(b)
    rel = C.c_double(0.0)
    it = lib().port_pcg_csr(len(b), rowptr.ctypes.data, col.ctypes.data, vals.ctypes.data, b.ctypes.data,
                            x.ctypes.data, tol, maxit, C.byref(rel), nthreads)
    return x, int(it), rel.value


def pcg_csr_batch(rowptr, col, vals, rhs, tol=1e-13, maxit=100000, nthreads=0):
    """Jacobi-PCG for many right-hand sides (B, n) on one matrix, OpenMP over the samples."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    x = np.empty_like(rhs)
    it = lib().port_pcg_csr_batch(rhs.shape[1], rowptr.ctypes.data, col.ctypes.data, vals.ctypes.data, rhs.shape[0],
                                  rhs.ctypes.data, x.ctypes.data, tol, maxit, nthreads)
    return x, int(it)
