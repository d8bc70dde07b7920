"""Synthetic workloads of BASELINE.json's configs.

SplitMix64-seeded uniform sample coordinates in the grid's bounding box (dimension-major:
all x, then y, then z), smooth-plus-noise values, unit-spaced grid at origin 0, sill 1,
nugget 0. The same arrays feed the oracle and the CUDA library.
"""
from __future__ import annotations

import math

import numpy as np

from . import _abi

_M64 = (1 << 64) - 1


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n successive SplitMix64 outputs as uint64 (vectorised: state_i = seed + (i+1)·γ)."""
    gamma = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _M64) + gamma * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(seed: int, n: int) -> np.ndarray:
    return (splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def make_samples(config_id: int, n: int, box, seed_extra: int = 0):
    """coords (list of dim arrays) and values for n samples in the box [0, box_d)."""
    dim = len(box)
    seed = (0x9E3779B97F4A7C15 ^ config_id ^ (seed_extra << 8)) & _M64
    u = uniform01(seed, n * (dim + 1))
    coords = [u[d * n:(d + 1) * n] * float(box[d]) for d in range(dim)]
    noise = u[dim * n:(dim + 1) * n]
    val = np.sin(2 * math.pi * coords[0] / box[0])
    if dim > 1:
        val = val + np.cos(2 * math.pi * coords[1] / box[1])
    if dim > 2:
        val = val + np.sin(2 * math.pi * coords[2] / box[2])
    val = val + 0.1 * (noise - 0.5)
    return coords, val


# name -> dict(dim, n, grid, estimator, vario, range, k, degree)
CONFIGS = {
    "C1": dict(id=1, n=500, grid=(100, 100), est="OK", vario="gaussian", range=35.0, k=0),
    "C2": dict(id=2, n=10_000, grid=(1000, 1000), est="OK", vario="spherical", range=50.0, k=20),
    "C3a": dict(id=3, n=100_000, grid=(256, 256, 256), est="SK", vario="exponential", range=30.0, k=32),
    "C3b": dict(id=3, n=100_000, grid=(256, 256, 256), est="UK", degree=1, vario="exponential", range=30.0, k=32),
    "C4": dict(id=4, n=20_000, grid=(2048, 2048), est="OK", vario="spherical", range=256.0, k=0),
    "C5": dict(id=5, n=1_000_000, grid=(512, 512, 512), est="OK", vario="spherical", range=40.0, k=64),
}
_VARIO = {"gaussian": _abi.VARIO_GAUSSIAN, "spherical": _abi.VARIO_SPHERICAL, "exponential": _abi.VARIO_EXPONENTIAL}
_EST = {"SK": _abi.EST_SIMPLE, "OK": _abi.EST_ORDINARY, "UK": _abi.EST_UNIVERSAL}


def config_spec(name: str, *, scale: float = 1.0, grid=None, n=None, k=None, support="block", ball_radius=None,
                min_neighbors=1, seed_extra=0) -> _abi.ProblemSpec:
    """ProblemSpec of a BASELINE config. `scale` < 1 shrinks the sample count and every grid axis
    (keeping the sample density) for parity tests that the oracle must finish in seconds."""
    cfg = dict(CONFIGS[name])
    g = tuple(grid) if grid is not None else tuple(max(4, int(round(s * scale))) for s in cfg["grid"])
    dim = len(g)
    if n is None:
        dens = cfg["n"] / float(np.prod(cfg["grid"]))
        n = cfg["n"] if scale == 1.0 and grid is None else max(cfg["k"] + 8, int(round(dens * float(np.prod(g)))))
    coords, vals = make_samples(cfg["id"], n, g, seed_extra)
    kk = cfg["k"] if k is None else k
    kw = dict(coords=coords, values=vals, grid_dims=g, grid_origin=[0.0] * dim, grid_spacing=[1.0] * dim,
              vario_kind=_VARIO[cfg["vario"]], vario_range=cfg["range"], vario_sill=1.0, vario_nugget=0.0,
              estimator=_EST[cfg["est"]], uk_degree=cfg.get("degree", 0), max_neighbors=min(kk, n),
              min_neighbors=min_neighbors)
    if cfg["est"] == "SK":
        kw["sk_mean"] = float(np.mean(vals))
    if support == "block":
        kw["support"] = _abi.default_support_py([1.0] * dim, cfg["range"])
    if ball_radius is not None:
        kw["ball_radius"] = float(ball_radius)
    return _abi.ProblemSpec(**kw)


def algorithmic_flops_per_target(spec: _abi.ProblemSpec) -> float:
    """SURVEY.md §8d's F_local / F_global (1 FMA = 2 flop; γ evaluation = 3d + C_γ)."""
    p = spec.params
    d = spec.dim
    cg = 10 if p["vario_kind"] == _abi.VARIO_SPHERICAL else 24
    per_eval = 3 * d + cg
    q = int(spec.support[0].shape[0])
    est = p["estimator"]
    c = 0 if est == _abi.EST_SIMPLE else (1 if est == _abi.EST_ORDINARY else math.comb(d + p["uk_degree"], d))
    k = p["max_neighbors"]
    if k > 0:
        m = k + c
        fact = m ** 3 / 3.0 if est == _abi.EST_SIMPLE else 2.0 * m ** 3 / 3.0
        return fact + 2 * m * m + 4 * m + (k * (k - 1) / 2 + k * q) * per_eval
    n = spec.n_samples
    return 2.0 * (n + c) ** 2 + n * q * per_eval + 4 * (n + c)
