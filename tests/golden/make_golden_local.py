"""Golden vectors for the LOCAL BASELINE configs AT FULL SAMPLE COUNT (C2: 1e4 samples, k = 20; C3a / C3b: 1e5 samples,
k = 32, SK / UK degree 1; C5: 1e6 samples, k = 64): 1 024 fixed-seed targets each, computed by the independent
numpy + LAPACK restatement (oracle/numpy_twin.py: brute-force k-NN with (d², index) ordering, dsytrf/dsytrs or
dpotrf/dpotrs — the factorisations Julia calls for the reference). Neighbour index lists, means and variances.

    python tests/golden/make_golden_local.py        ->  tests/golden/local_full_1024.npz   (about two minutes)
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy_twin as TW  # noqa: E402
import gskrige  # noqa: E402

NT = 1024
EST = {gskrige.EST_SIMPLE: TW.SIMPLE, gskrige.EST_ORDINARY: TW.ORDINARY, gskrige.EST_UNIVERSAL: TW.UNIVERSAL}


def main():
    out = {}
    for name in ("C2", "C3a", "C3b", "C5"):
        t0 = time.time()
        spec = gskrige.synth.config_spec(name)
        T = spec.n_targets
        targets = np.sort(np.random.default_rng(abs(hash(name)) % 2**31 if False else {"C2": 2, "C3a": 31, "C3b": 32, "C5": 5}[name])
                          .choice(T, NT, replace=False)).astype(np.int64)
        lin, ctr = targets.copy(), []
        for d in range(spec.dim):
            ctr.append(spec.grid_origin[d] + ((lin % spec.grid_dims[d]).astype(np.float64) + 0.5) * spec.grid_spacing[d])
            lin //= spec.grid_dims[d]
        p = spec.params
        vario = dict(kind=p["vario_kind"], range=p["vario_range"], sill=p["vario_sill"], nugget=p["vario_nugget"], eps=p["gaussian_nugget_eps"])
        mean, var, nn, idx = TW.krige(np.stack(spec.coords, 1), spec.values, np.stack(ctr, 1), support=np.stack(spec.support, 1),
                                      vario=vario, est=EST[p["estimator"]], sk_mean=p["sk_mean"], degree=p["uk_degree"],
                                      k=p["max_neighbors"], min_neighbors=p["min_neighbors"])
        out[f"{name}/targets"], out[f"{name}/mean"], out[f"{name}/var"], out[f"{name}/idx"] = targets, mean, var, idx.astype(np.int32)
        print(f"{name}: {NT} targets in {time.time() - t0:.1f} s", flush=True)
    np.savez_compressed(Path(__file__).parent / "local_full_1024.npz", **out)


if __name__ == "__main__":
    main()
