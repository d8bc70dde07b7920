"""ctypes wrapper of the CPU oracle (oracle/gsk_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
LIB = ORACLE_DIR / "_build" / "libgskoracle.so"
SEARCH_BRUTE, SEARCH_KDTREE = 0, 1

_lib = None


def build(force: bool = False) -> Path:
    src = ORACLE_DIR / "gsk_oracle.c"
    hdr = ORACLE_DIR.parent / "include" / "gskrige.h"
    if force or not LIB.exists() or LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-s", "-B"], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB))
        vp = C.c_void_p
        lib.gsk_oracle_krige.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int]
        lib.gsk_oracle_krige.restype = C.c_int
        lib.gsk_oracle_search.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int]
        lib.gsk_oracle_search.restype = C.c_int
        lib.gsk_oracle_uk_exponents.argtypes = [C.c_int, C.c_int, vp, C.c_int]
        lib.gsk_oracle_uk_exponents.restype = C.c_int
        lib.gsk_oracle_threads.restype = C.c_int
        lib.gsk_oracle_sgs.argtypes = ([C.c_int, C.c_int64, vp, vp, C.c_int] + [C.c_double] * 5 + [C.c_int, C.c_int, C.c_double]
                                       + [vp] * 7)
        lib.gsk_oracle_sgs.restype = C.c_int
        _lib = lib
    return _lib


def krige(spec, search=SEARCH_KDTREE, nthreads=0, want_neighbors=False):
    """spec: geostatssolvers.jl_b200 ProblemSpec (same bytes the CUDA library receives)."""
    lib = load()
    first, count = spec.slab
    mean = np.empty(count, dtype=np.float64)
    var = np.empty(count, dtype=np.float64)
    k = spec.params["max_neighbors"]
    nneigh = np.empty(count, dtype=np.int32)
    idx = np.empty((count, max(k, 1)), dtype=np.int32) if k > 0 else None
    ps = spec.c_struct()
    rc = lib.gsk_oracle_krige(C.addressof(ps), mean.ctypes.data, var.ctypes.data, nneigh.ctypes.data,
                              idx.ctypes.data if idx is not None else None, search, nthreads)
    if rc != 0:
        raise RuntimeError(f"gsk_oracle_krige failed: {rc}")
    if want_neighbors:
        return mean, var, nneigh, idx
    return mean, var


def search(spec, search=SEARCH_BRUTE, nthreads=0):
    lib = load()
    first, count = spec.slab
    k = spec.params["max_neighbors"]
    nneigh = np.empty(count, dtype=np.int32)
    idx = np.empty((count, k), dtype=np.int32)
    d2 = np.empty((count, k), dtype=np.float64)
    ps = spec.c_struct()
    rc = lib.gsk_oracle_search(C.addressof(ps), nneigh.ctypes.data, idx.ctypes.data, d2.ctypes.data, search, nthreads)
    if rc != 0:
        raise RuntimeError(f"gsk_oracle_search failed: {rc}")
    return nneigh, idx, d2


def sgs(coords, rank, *, vario_kind, vario_range, vario_sill=1.0, vario_nugget=0.0, gaussian_nugget_eps=1e-6, mean=0.0,
        min_neighbors=1, max_neighbors=10, ball_radius=float("nan"), values=None, z=None):
    """One realisation of the sequential loop (seq.jl:102-135); returns (out, nneigh, idx, weights, sigma)."""
    lib = load()
    cs = [np.ascontiguousarray(c, dtype=np.float64) for c in coords]
    n = cs[0].shape[0]
    k = min(max_neighbors, n)
    arr = (C.c_void_p * 3)(*[c.ctypes.data for c in cs] + [None] * (3 - len(cs)))
    rank = np.ascontiguousarray(rank, dtype=np.int64)
    vals = np.ascontiguousarray(values, dtype=np.float64) if values is not None else None
    z = np.ascontiguousarray(z, dtype=np.float64)
    out = np.empty(n)
    nn = np.zeros(n, dtype=np.int32)
    idx = np.full((n, k), -1, dtype=np.int32)
    lam = np.zeros((n, k))
    sig = np.zeros(n)
    rc = lib.gsk_oracle_sgs(len(cs), n, C.addressof(arr), rank.ctypes.data, int(vario_kind), float(vario_range),
                            float(vario_sill), float(vario_nugget), float(gaussian_nugget_eps), float(mean),
                            int(min_neighbors), int(max_neighbors), float(ball_radius),
                            vals.ctypes.data if vals is not None else None, z.ctypes.data, out.ctypes.data,
                            nn.ctypes.data, idx.ctypes.data, lam.ctypes.data, sig.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"gsk_oracle_sgs failed: {rc}")
    return out, nn, idx, lam, sig


def uk_exponents(degree, dim):
    lib = load()
    out = np.zeros((16, dim), dtype=np.int32)
    n = lib.gsk_oracle_uk_exponents(degree, dim, out.ctypes.data, 16)
    if n < 0:
        raise ValueError("bad degree/dim")
    return out[:n].copy()


def threads():
    return load().gsk_oracle_threads()
