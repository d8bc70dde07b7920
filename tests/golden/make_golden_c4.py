"""Golden vectors for BASELINE config C4 AT FULL SIZE (global Ordinary Kriging, 20 000 2-D samples, SphericalVariogram
(range = 256), 2048 x 2048 grid): 4 096 fixed-seed targets solved with the factorisation the reference itself would
use — LAPACK dsytrf/dsytrs (Bunch–Kaufman, upper), i.e. Julia's bunchkaufman(Symmetric(LHS)) \\ RHS on the
20 001 x 20 001 kriging system (GeoStatsModels 0.2 [3P]) — through numpy/scipy, independently of both
the C oracle (partial-pivot LU) and the CUDA library (Cholesky + Schur complement on L^-1).

The reference cannot run here (no julia binary); this is the LAPACK restatement at the full problem size.
~3.3 GB of RAM and about two minutes on 8 cores:

    python tests/golden/make_golden_c4.py          ->  tests/golden/c4_full_4096.npz
"""
import sys
import time
from pathlib import Path

import numpy as np
from scipy.linalg import lapack

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy_twin as TW  # noqa: E402
import gskrige  # noqa: E402

NT = 4096


def main():
    spec = gskrige.synth.config_spec("C4")
    n = spec.n_samples
    X = np.stack(spec.coords, 1)
    p = spec.params
    sill, rng, nug = p["vario_sill"], p["vario_range"], p["vario_nugget"]
    m = n + 1
    t0 = time.time()
    A = np.zeros((m, m), order="F")
    for lo in range(0, n, 2000):                      # covariance block, 2000 columns at a time
        hi = min(n, lo + 2000)
        dx = X[:, None, 0] - X[None, lo:hi, 0]
        dy = X[:, None, 1] - X[None, lo:hi, 1]
        h = np.sqrt(dx * dx + dy * dy)
        A[:n, lo:hi] = sill - TW.variogram(TW.SPHERICAL, h, rng, sill, nug)
    A[n, :n] = 1.0
    A[:n, n] = 1.0
    print(f"assembled {m}x{m} in {time.time() - t0:.1f} s", flush=True)
    t0 = time.time()
    ldu, ipiv, info = lapack.dsytrf(A, lower=0, overwrite_a=1)
    assert info == 0
    print(f"dsytrf in {time.time() - t0:.1f} s", flush=True)
    T = spec.n_targets
    targets = np.sort(np.random.default_rng(20241018).choice(T, NT, replace=False)).astype(np.int64)
    gx, gy = spec.grid_dims
    cx = spec.grid_origin[0] + ((targets % gx).astype(np.float64) + 0.5) * spec.grid_spacing[0]
    cy = spec.grid_origin[1] + ((targets // gx).astype(np.float64) + 0.5) * spec.grid_spacing[1]
    sup = np.stack(spec.support, 1)
    B = np.zeros((m, NT), order="F")
    acc = np.zeros((n, NT))
    for s in range(sup.shape[0]):                     # RHS: mean of gamma over the block-support points
        dx = (cx[None, :] + sup[s, 0]) - X[:, None, 0]
        dy = (cy[None, :] + sup[s, 1]) - X[:, None, 1]
        acc += TW.variogram(TW.SPHERICAL, np.sqrt(dx * dx + dy * dy), rng, sill, nug)
    B[:n] = sill - acc / sup.shape[0]
    B[n] = 1.0
    t0 = time.time()
    S, info = lapack.dsytrs(ldu, ipiv, B, lower=0)
    assert info == 0
    print(f"dsytrs ({NT} right-hand sides) in {time.time() - t0:.1f} s", flush=True)
    lam = S[:n]
    mean = lam.T @ spec.values
    var = sill - (np.einsum("ij,ij->j", B[:n], lam) + B[n] * S[n])
    var = np.sqrt(np.maximum(var, 0.0)) ** 2          # predictvar clamp + Normal(mu, sqrt(var)) round trip (krig.jl:183)
    np.savez_compressed(Path(__file__).parent / "c4_full_4096.npz", targets=targets, mean=mean, var=var)
    print("wrote c4_full_4096.npz", float(mean.min()), float(mean.max()), float(var.min()), float(var.max()))


if __name__ == "__main__":
    main()
