"""Generates tests/golden/kriging_small.npz with oracle/numpy_twin.py (numpy + LAPACK dsytrf/dpotrf —
the factorisations Julia's bunchkaufman/cholesky call for the reference).

The reference itself cannot run here (no julia binary; its arithmetic lives in un-vendored packages),
so these vectors come from the independent LAPACK restatement, plus the reference's own nine
known-answer checks (test/estimation/krig.jl:35-37,50-52,70-72) stored as `ref_checks`.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy_twin as TW  # noqa: E402
import gskrige  # noqa: E402

CASES = [
    # name, dim, n, grid, est, degree, vario kind, range, k, radius
    ("ok_gauss_2d_global", 2, 60, (12, 12), 1, 0, 0, 6.0, 0, None),
    ("ok_sph_2d_k8", 2, 120, (16, 16), 1, 0, 1, 9.0, 8, None),
    ("sk_exp_3d_k12", 3, 300, (8, 8, 8), 0, 0, 2, 6.0, 12, None),
    ("uk1_exp_3d_k16", 3, 300, (8, 8, 8), 2, 1, 2, 6.0, 16, None),
    ("uk2_sph_2d_k20", 2, 200, (14, 14), 2, 2, 1, 10.0, 20, None),
    ("ok_gauss_1d_ball", 1, 40, (50,), 1, 0, 0, 8.0, 5, 6.0),
    ("ok_sph_2d_ball_min4", 2, 100, (20, 20), 1, 0, 1, 8.0, 10, 3.0),
]


def build(case):
    name, dim, n, grid, est, degree, vk, rng, k, radius = case
    coords, vals = gskrige.synth.make_samples(77, n, grid, seed_extra=len(name))
    sup = gskrige.default_support_py([1.0] * dim, rng)
    spec = gskrige.ProblemSpec(coords=coords, values=vals, grid_dims=grid, support=sup, vario_kind=vk, vario_range=rng,
                               estimator=est, uk_degree=degree, sk_mean=float(np.mean(vals)), max_neighbors=k,
                               ball_radius=float("nan") if radius is None else radius,
                               min_neighbors=4 if "min4" in name else 1)
    return spec


def main():
    out = {}
    for case in CASES:
        name, dim, n, grid, est, degree, vk, rng, k, radius = case
        spec = build(case)
        X = np.stack(spec.coords, 1)
        ctr = np.stack(spec.target_centers(), 1)
        sup = np.stack(spec.support, 1)
        vario = dict(kind=vk, range=rng, sill=1.0, nugget=0.0, eps=1e-6)
        mean, var, nn, idx = TW.krige(X, spec.values, ctr, support=sup, vario=vario, est=est, sk_mean=spec.params["sk_mean"],
                                      degree=degree, k=k, radius=radius, min_neighbors=spec.params["min_neighbors"])
        out[f"{name}/mean"] = mean
        out[f"{name}/var"] = var
        out[f"{name}/nn"] = nn
        if k > 0:
            out[f"{name}/idx"] = idx
    # the reference's own known answers: (i, j) 1-based grid index -> expected value (atol 1e-3)
    out["ref_checks"] = np.array([[25, 25, 1.0], [50, 75, 0.0], [75, 50, 1.0]])
    np.savez_compressed(Path(__file__).parent / "kriging_small.npz", **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
