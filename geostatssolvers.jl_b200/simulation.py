"""Host-side mirror of the reference's FFT Gaussian simulation solver — the first CALLER of the Kriging hot path
(SURVEY §8f-2): ref src/simulation/fft.jl.

    preprocess  fft.jl:62-135   spectrum of the covariance (one FFT), and for CONDITIONAL simulation one Simple Kriging
                                solve of the data onto `PointSet(centroid.(pdomain))` (fft.jl:112-126)
    solvesingle fft.jl:145-192  per realisation: random phases → inverse FFT → rescale; conditioning solves THE SAME
                                Simple Kriging problem again with the unconditional realisation's values at the data
                                cells (fft.jl:175-188) and adds the residual field

What runs where: the FFTs go through torch.fft (cuFFT — a library call; they are not the hot path this repository
accelerates), the Kriging solves through libgskrige.so with GSK_FLAG_REUSE_PLAN — the per-realisation solves share the
sample coordinates, so only the VALUES cross the ABI (gsk_update_values inside gsk_krige): bins / neighbour lists
(local) or L and L^-1 (global) stay resident in HBM between realisations.

Julia's RNG cannot be reproduced here: `rng` is a numpy Generator (or a seed); the uniform noise field of a
realisation can also be passed explicitly (`noise=`), which is what the parity tests do.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _abi
from .host import (CartesianGrid, Euclidean, GaussianVariogram, GeoTable, KBallSearch, PointSet, UnsupportedOption, _Variogram,
                   _unsupported, default_context, georef, searcher_ui)


class SimulationProblem:
    """``SimulationProblem(domain, "z", nreals)`` or ``SimulationProblem(samples, domain, "z", nreals)``
    (GeoStatsBase; ref test/simulation/fft.jl:4,33)."""

    def __init__(self, *args):
        if isinstance(args[0], GeoTable):
            self._data, self._domain, self._var, self._nreals = args
        else:
            self._data = None
            self._domain, self._var, self._nreals = args
        self._var = self._var if isinstance(self._var, str) else self._var[0]

    def data(self):
        return self._data

    def domain(self):
        return self._domain

    def variables(self):
        return (self._var,)

    def nreals(self):
        return int(self._nreals)


_FFTGS_DEFAULTS = dict(variogram=None, mean=0.0, minneighbors=1, maxneighbors=None, neighborhood=None, distance=None)


class FFTGS:
    """``FFTGS(z=dict(variogram=GaussianVariogram(range=10.0)), rng=2019)`` — parameters: ref fft.jl:50-59."""

    def __init__(self, *pairs, rng=None, **kwpairs):
        self.vparams = {}
        for item in list(pairs) + list(kwpairs.items()):
            for var, params in (list(item.items()) if isinstance(item, dict) else [item]):
                unknown = set(params) - set(_FFTGS_DEFAULTS)
                if unknown:
                    raise TypeError(f"unknown FFTGS parameter(s) {sorted(unknown)} for variable {var}")
                self.vparams[var] = dict(params)
        self.rng = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)

    def params(self, var):
        p = dict(_FFTGS_DEFAULTS)
        p.update(self.vparams.get(var, {}))
        if p["variogram"] is None:
            p["variogram"] = GaussianVariogram()
        if p["distance"] is None:
            p["distance"] = Euclidean()
        return p


def variogram_values(gamma: _Variogram, h: np.ndarray) -> np.ndarray:
    """γ(h), point to point (Variography formulas) — host side, for the spectrum only."""
    h = np.asarray(h, dtype=np.float64)
    s, r = gamma.sill, gamma.range
    n = gamma.nugget + (1e-6 if gamma.kind == _abi.VARIO_GAUSSIAN else 0.0)
    if gamma.kind == _abi.VARIO_GAUSSIAN:
        g = (s - n) * (1.0 - np.exp(-3.0 * (h / r) ** 2))
    elif gamma.kind == _abi.VARIO_SPHERICAL:
        t = h / r
        g = np.where(h < r, (s - n) * (1.5 * t - 0.5 * t ** 3), (s - n))
    else:
        g = (s - n) * (1.0 - np.exp(-3.0 * (h / r)))
    return g + np.where(h > 0, n, 0.0)


def _fft_device():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("FFTGS runs its FFTs on the GPU (torch.fft / cuFFT); no CUDA device is visible")
    return torch, torch.device("cuda", torch.cuda.current_device())


def _kriging_spec(gamma, mean, p, data_coords, data_vals, target_pts):
    """SK of the data onto PointSet(centroid.(pdomain)) — the KrigingSolver of fft.jl:115-124 as a ProblemSpec."""
    if not isinstance(p["distance"], Euclidean):
        raise _unsupported("non-Euclidean `distance`")
    n = data_vals.shape[0]
    dim = len(data_coords)
    sdom = PointSet(np.stack(data_coords, 0))
    kw = dict(coords=data_coords, values=data_vals, points=target_pts, vario_kind=gamma.kind, vario_range=gamma.range,
              vario_sill=gamma.sill, vario_nugget=gamma.nugget, estimator=_abi.EST_SIMPLE, sk_mean=float(mean),
              min_neighbors=int(p["minneighbors"]), flags=_abi.FLAGS_DEFAULT | _abi.FLAG_REUSE_PLAN)
    if p["maxneighbors"] is not None:                                  # krig.jl:151: local iff maxneighbors given
        searcher = searcher_ui(sdom, p["maxneighbors"], p["distance"], p["neighborhood"])
        if searcher.k > _abi.GSK_MAX_NEIGHBORS:
            raise _unsupported(f"maxneighbors > {_abi.GSK_MAX_NEIGHBORS} on the local path")
        kw.update(max_neighbors=searcher.k)
        if isinstance(searcher, KBallSearch):
            kw.update(ball_radius=searcher.ball.radius())
    del n, dim
    return _abi.ProblemSpec(**kw)


def preprocess_fftgs(problem: SimulationProblem, solver: FFTGS, ctx: Optional[_abi.Context] = None) -> dict:
    """ref fft.jl:62-135"""
    pgrid = problem.domain()
    if not isinstance(pgrid, CartesianGrid):
        raise _unsupported("FFTGS on a domain that is not a CartesianGrid")
    torch, dev = _fft_device()
    dims = pgrid.dims
    var = problem.variables()[0]
    p = solver.params(var)
    gamma, mu = p["variogram"], float(p["mean"])
    if not isinstance(gamma, _Variogram) or gamma.kind < 0:
        raise _unsupported(f"variogram {type(gamma).__name__}")
    cents = pgrid.centroids()                                          # x fastest
    cidx = [d // 2 - 1 for d in dims]                                  # CartesianIndex(dims .÷ 2), 0-based (fft.jl:69)
    ccen = [pgrid.origin[a] + (max(cidx[a], 0) + 0.5) * pgrid.spacing[a] for a in range(len(dims))]
    h = np.sqrt(sum((cents[a] - ccen[a]) ** 2 for a in range(len(dims))))
    cov = gamma.sill - variogram_values(gamma, h)                      # fft.jl:98-100
    C = torch.from_numpy(np.reshape(cov, dims, order="F").copy()).to(dev)
    F = torch.sqrt(torch.abs(torch.fft.fftn(torch.fft.fftshift(C))))   # fft.jl:103
    F.view(-1)[0] = 0.0                                                # fft.jl:104: F[1] = 0 (Julia's first element)
    pre = dict(gamma=gamma, mu=mu, F=F, zbar=None, spec=None, dinds=None, p=p, dims=dims, cents=cents)
    pdata = problem.data()
    if pdata is not None and var in pdata.table:                       # fft.jl:108-134
        ddom = pdata.domain
        if not isinstance(ddom, PointSet):
            raise _unsupported("sample domains that are not point sets")
        dcoords = ddom.centroids()
        dvals = np.asarray(pdata.table[var], dtype=np.float64)
        spec = _kriging_spec(gamma, mu, p, dcoords, dvals, cents)
        zbar, _ = (ctx or default_context()).krige(spec)               # fft.jl:125-126
        # nearest grid element of every datum (KNearestSearch(pdomain, 1), fft.jl:129-133); unique, first occurrence
        ijk = [np.clip(np.floor((dcoords[a] - pgrid.origin[a]) / pgrid.spacing[a]).astype(np.int64), 0, dims[a] - 1)
               for a in range(len(dims))]
        lin = np.zeros_like(ijk[0])
        for a in reversed(range(len(dims))):
            lin = lin * dims[a] + ijk[a]
        _, first = np.unique(lin, return_index=True)
        pre.update(zbar=zbar, dinds=lin[np.sort(first)])
    return pre


def solvesingle_fftgs(problem: SimulationProblem, solver: FFTGS, pre: dict, ctx: Optional[_abi.Context] = None,
                      noise: Optional[np.ndarray] = None) -> np.ndarray:
    """ref fft.jl:145-192 — one realisation (a flat array in domain order)."""
    torch, dev = _fft_device()
    dims, gamma, mu, F = pre["dims"], pre["gamma"], pre["mu"], pre["F"]
    if noise is None:
        noise = solver.rng.random(dims[::-1]).transpose()              # rand(rng, V, dims), column-major fill
    U = torch.from_numpy(np.ascontiguousarray(np.reshape(np.asarray(noise, dtype=np.float64), dims))).to(dev)
    P = F * torch.exp(1j * torch.angle(torch.fft.fftn(U)))             # fft.jl:161
    Z = torch.real(torch.fft.ifftn(P))                                 # fft.jl:164
    s2 = torch.mean(Z * Z) * (Z.numel() / (Z.numel() - 1))             # Statistics.var(Z, mean=0): corrected (fft.jl:167)
    Z = torch.sqrt(gamma.sill / s2) * Z + mu                           # fft.jl:168
    zu = Z.cpu().numpy().reshape(-1, order="F")                        # Z[inds], x fastest
    if pre["zbar"] is None:
        return zu
    dinds = pre["dinds"]
    cents = pre["cents"]
    spec = _kriging_spec(gamma, mu, pre["p"], [c[dinds] for c in cents], zu[dinds].copy(), cents)   # fft.jl:176-186
    zbar_u, _ = (ctx or default_context()).krige(spec)                 # same coordinates every realisation: values-only update
    return pre["zbar"] + (zu - zbar_u)                                 # fft.jl:189


def solve_fftgs(problem: SimulationProblem, solver: FFTGS, ctx: Optional[_abi.Context] = None):
    """``solve(problem, FFTGS(...))`` — a list of `nreals` GeoTables (the reference returns an Ensemble)."""
    pre = preprocess_fftgs(problem, solver, ctx)
    var = problem.variables()[0]
    return [georef({var: solvesingle_fftgs(problem, solver, pre, ctx)}, problem.domain()) for _ in range(problem.nreals())]


# --------------------------------------------------------------------------------------------------------------
# LUGS — LU Gaussian simulation: ref src/simulation/lu.jl
#   preprocess  lu.jl:75-160   data cells (NearestInit), covariance blocks, Cholesky factors, conditional mean d2
#   lusim       lu.jl:198-224  y2 = d2 + L22·w2 (second variable: w = ρ·w1 + sqrt(1−ρ²)·w2)
# The dense factorisation and the matrix–vector product of every realisation run in libgskrige.so (gsk_lu_plan /
# gsk_lu_sample: the blocked FP64 Cholesky of the global Kriging plan on the JOINT covariance [C11 C12; C21 C22]).
# --------------------------------------------------------------------------------------------------------------
_LUGS_DEFAULTS = dict(variogram=None, mean=None, factorization="cholesky")


class LUGS:
    """``LUGS(z=dict(variogram=SphericalVariogram(range=10.0)), rng=2019)`` — parameters: ref lu.jl:66-73. Only the
    default `factorization=cholesky` exists on the GPU (`lu` is rejected); `correlation` couples two variables."""

    def __init__(self, *pairs, rng=None, correlation=0.0, **kwpairs):
        self.vparams = {}
        for item in list(pairs) + list(kwpairs.items()):
            for var, params in (list(item.items()) if isinstance(item, dict) else [item]):
                unknown = set(params) - set(_LUGS_DEFAULTS)
                if unknown:
                    raise TypeError(f"unknown LUGS parameter(s) {sorted(unknown)} for variable {var}")
                self.vparams[var] = dict(params)
        self.correlation = float(correlation)
        self.rng = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)

    def params(self, var):
        p = dict(_LUGS_DEFAULTS)
        p.update(self.vparams.get(var, {}))
        if p["variogram"] is None:
            p["variogram"] = GaussianVariogram()
        if p["factorization"] not in ("cholesky", None):
            raise _unsupported("LUGS `factorization` other than cholesky")
        return p


def _nearest_cells(pdomain, coords):
    """NearestInit (GeoStatsBase initbuff): every datum goes to the domain element nearest to it; a later datum
    overwrites an earlier one in the same element. Returns (cells in first-seen order, index of the datum kept)."""
    cents = pdomain.centroids()
    n = coords[0].shape[0]
    if isinstance(pdomain, CartesianGrid):
        dims = pdomain.dims
        ijk = [np.clip(np.floor((coords[a] - pdomain.origin[a]) / pdomain.spacing[a]).astype(np.int64), 0, dims[a] - 1)
               for a in range(len(dims))]
        lin = np.zeros(n, dtype=np.int64)
        for a in reversed(range(len(dims))):
            lin = lin * dims[a] + ijk[a]
    else:
        P = np.stack(cents, 1)
        lin = np.array([int(np.argmin(((P - np.array([c[i] for c in coords])) ** 2).sum(1))) for i in range(n)], dtype=np.int64)
    keep = {}
    for i, c in enumerate(lin.tolist()):
        keep[c] = i
    cells = np.array(sorted(keep), dtype=np.int64)               # findall(mask): ascending element index (lu.jl:113)
    return cells, np.array([keep[c] for c in cells.tolist()], dtype=np.int64)


def preprocess_lugs(problem: SimulationProblem, solver: LUGS, var: str, ctx: _abi.Context) -> dict:
    """ref lu.jl:75-160 for one variable: the joint covariance is factorised on the GPU and stays resident in `ctx`"""
    pdomain = problem.domain()
    p = solver.params(var)
    gamma = p["variogram"]
    if not isinstance(gamma, _Variogram) or gamma.kind < 0:
        raise _unsupported(f"variogram {type(gamma).__name__}")
    cents = pdomain.centroids()
    npts = pdomain.nelements()
    pdata = problem.data()
    if pdata is not None and var in pdata.table:
        dcoords = pdata.domain.centroids()
        dlocs, kept = _nearest_cells(pdomain, dcoords)
        z1 = np.asarray(pdata.table[var], dtype=np.float64)[kept]
    else:
        dlocs, z1 = np.zeros(0, dtype=np.int64), np.zeros(0)
    mask = np.ones(npts, dtype=bool)
    mask[dlocs] = False
    slocs = np.flatnonzero(mask)                                  # lu.jl:117
    order = np.concatenate([dlocs, slocs])
    ctx.lu_plan([c[order] for c in cents], len(dlocs), z1, vario_kind=gamma.kind, vario_range=gamma.range,
                vario_sill=gamma.sill, vario_nugget=gamma.nugget)
    if p["mean"] is not None and len(dlocs) > 0:
        import warnings
        warnings.warn("mean can only be specified in unconditional simulation")   # lu.jl:137-139
    mu = 0.0 if p["mean"] is None else float(p["mean"])
    return dict(dlocs=dlocs, slocs=slocs, z1=z1, mu=mu, npts=npts)


def lusim(ctx: _abi.Context, pre: dict, w2: np.ndarray, rho=None, w1=None) -> np.ndarray:
    """ref lu.jl:198-224 for the draws w2 (and, for the second of two correlated variables, ρ and the first's draws)"""
    w = w2 if rho is None else rho * w1 + np.sqrt(1.0 - rho ** 2) * w2
    joint = ctx.lu_sample(w)                                       # [z1; d2 + L22·w]
    nd = len(pre["dlocs"])
    y = np.empty(pre["npts"])
    y[pre["dlocs"]] = joint[:nd]
    y[pre["slocs"]] = joint[nd:]
    if nd == 0:
        y += pre["mu"]                                             # lu.jl:221
    return y


def solve_lugs(problem: SimulationProblem, solver: LUGS, ctx: Optional[_abi.Context] = None):
    """``solve(problem, LUGS(...))`` for one variable — a list of `nreals` GeoTables."""
    ctx = ctx or default_context()
    var = problem.variables()[0]
    pre = preprocess_lugs(problem, solver, var, ctx)
    ns = len(pre["slocs"])
    return [georef({var: lusim(ctx, pre, solver.rng.standard_normal(ns))}, problem.domain()) for _ in range(problem.nreals())]


# --------------------------------------------------------------------------------------------------------------
# Sequential Gaussian simulation — ref src/simulation/sgs.jl (parameters, estimator, marginal) and
# src/simulation/seq.jl (the loop).
#   preprocess  sgs.jl:56-89 → seq.jl:42-74   SimpleKriging(variogram, mean), Normal(mean, √sill), searcher_ui
#   solvesingle seq.jl:76-141                 for ind in traverse(pdomain, path): masked search → fit → draw
# The masked search, the fit and the kriging weights of EVERY location do not depend on the simulated values; they
# run once, in parallel, in libgskrige.so (gsk_sgs_plan) and stay resident. A realisation is then the cheap
# recurrence over the path (gsk_sgs_sample), and all `nreals` realisations run at the same time.
# --------------------------------------------------------------------------------------------------------------
_SGS_DEFAULTS = dict(variogram=None, mean=0.0, path=None, minneighbors=1, maxneighbors=10, neighborhood=None, distance=None)
SGS_MAX_NEIGHBORS = 64


class SGS:
    """``SGS(z=dict(variogram=SphericalVariogram(range=35.0), neighborhood=MetricBall(10.0)), rng=2017)`` —
    parameters: ref sgs.jl:44-54. `init` is NearestInit (the only one the reference's tests use)."""

    def __init__(self, *pairs, rng=None, **kwpairs):
        self.vparams = {}
        for item in list(pairs) + list(kwpairs.items()):
            for var, params in (list(item.items()) if isinstance(item, dict) else [item]):
                unknown = set(params) - set(_SGS_DEFAULTS)
                if unknown:
                    raise TypeError(f"unknown SGS parameter(s) {sorted(unknown)} for variable {var}")
                self.vparams[var] = dict(params)
        self.rng = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)

    def params(self, var):
        p = dict(_SGS_DEFAULTS)
        p.update(self.vparams.get(var, {}))
        if p["variogram"] is None:
            p["variogram"] = GaussianVariogram()
        if p["distance"] is None:
            p["distance"] = Euclidean()
        if p["path"] is None:
            from .host import LinearPath
            p["path"] = LinearPath()
        return p


def preprocess_sgs(problem: SimulationProblem, solver: SGS, var: str, ctx: _abi.Context) -> dict:
    """ref sgs.jl:56-89 + seq.jl:42-74 for one variable, plus everything of the loop (seq.jl:102-127) that does not
    depend on the simulated values: the plan is resident in `ctx` afterwards"""
    from .host import traverse
    pdomain = problem.domain()
    p = solver.params(var)
    gamma = p["variogram"]
    if not isinstance(gamma, _Variogram) or gamma.kind < 0:
        raise _unsupported(f"variogram {type(gamma).__name__}")
    if not isinstance(p["distance"], Euclidean):
        raise _unsupported("non-Euclidean `distance`")
    npts = pdomain.nelements()
    searcher = searcher_ui(pdomain, p["maxneighbors"], p["distance"], p["neighborhood"])   # seq.jl:66
    if searcher.k > SGS_MAX_NEIGHBORS:
        raise _unsupported(f"maxneighbors > {SGS_MAX_NEIGHBORS} in sequential simulation")
    radius = searcher.ball.radius() if isinstance(searcher, KBallSearch) else float("nan")
    # initbuff(pdomain, pvars, NearestInit(), data=pdata): buffer and mask (seq.jl:88)
    values = np.zeros(npts)
    mask = np.zeros(npts, dtype=bool)
    pdata = problem.data()
    if pdata is not None and var in pdata.table:
        cells, kept = _nearest_cells(pdomain, pdata.domain.centroids())
        col = np.asarray(pdata.table[var], dtype=np.float64)[kept]
        good = ~np.isnan(col)                                       # missing values are not copied into the buffer
        values[cells[good]] = col[good]
        mask[cells[good]] = True
    # traverse(pdomain, path), skipping the locations that hold data (seq.jl:102-103)
    visit = traverse(pdomain, p["path"])
    visit = np.arange(npts, dtype=np.int64) if visit is None else np.asarray(visit, dtype=np.int64)
    visit = visit[~mask[visit]]
    rank = np.full(npts, -1, dtype=np.int64)
    rank[visit] = np.arange(len(visit), dtype=np.int64)
    ctx.sgs_plan(pdomain.centroids(), rank, vario_kind=gamma.kind, vario_range=gamma.range, vario_sill=gamma.sill,
                 vario_nugget=gamma.nugget, mean=float(p["mean"]), min_neighbors=int(p["minneighbors"]),
                 max_neighbors=int(searcher.k), ball_radius=radius)
    return dict(values=values, mask=mask, visit=visit, rank=rank, npts=npts)


def sgs_draws(pre: dict, rng: np.random.Generator, nreals: int) -> np.ndarray:
    """the standard normal draws of `nreals` realisations, laid out per location: the p-th draw of a realisation belongs
    to the p-th visited location (seq.jl:110,130 draw one number per visited location, in path order)"""
    z = np.zeros((nreals, pre["npts"]))
    for r in range(nreals):
        z[r, pre["visit"]] = rng.standard_normal(len(pre["visit"]))
    return z


def solve_sgs(problem: SimulationProblem, solver: SGS, ctx: Optional[_abi.Context] = None, z: Optional[np.ndarray] = None,
              device_draws: bool = False):
    """``solve(problem, SGS(...))`` for one variable — a list of `nreals` GeoTables. `z` (nreals × nelements) overrides
    the draws (parity tests hand the oracle the same numbers). `device_draws=True` draws on the GPU (torch.randn with a
    generator seeded from `solver.rng`) and samples on device buffers (gsk_sgs_sample_device): only the realisations
    cross PCIe, once, through pinned memory."""
    ctx = ctx or default_context()
    var = problem.variables()[0]
    pre = preprocess_sgs(problem, solver, var, ctx)
    nreals = problem.nreals()
    if device_draws and z is None:
        import torch
        dev = torch.device("cuda", ctx.device)
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(solver.rng.integers(0, 2 ** 62)))
        dz = torch.randn((nreals, pre["npts"]), dtype=torch.float64, device=dev, generator=gen)
        dvals = torch.from_numpy(pre["values"]).to(dev)
        dout = torch.empty_like(dz)
        torch.cuda.synchronize(dev)
        ctx.sgs_sample_device(nreals, dvals.data_ptr(), dz.data_ptr(), dout.data_ptr())
        ctx.synchronize()
        host = torch.empty(dout.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(dout)
        reals = host.numpy()
        return [georef({var: reals[r]}, problem.domain()) for r in range(nreals)]
    if z is None:
        z = sgs_draws(pre, solver.rng, nreals)
    reals = ctx.sgs_sample(np.asarray(z, dtype=np.float64).reshape(nreals, pre["npts"]), values=pre["values"])
    return [georef({var: reals[r]}, problem.domain()) for r in range(nreals)]
