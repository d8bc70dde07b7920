"""Ad-hoc GPU parity check against the CPU oracle (development aid; the real tests are in tests/)."""
import sys, time, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np
import gskrige
import oracle_py as O

def compare(name, spec, search=O.SEARCH_KDTREE):
    ctx = gskrige.default_context()
    t0 = time.time()
    out = ctx.krige(spec, want_neighbors=True)
    t1 = time.time()
    mean, var, nn, idx = out
    om, ov, onn, oidx = O.krige(spec, search=search, want_neighbors=True)
    t2 = time.time()
    ok_nn = np.array_equal(nn, onn)
    ok_idx = True if idx is None else np.array_equal(idx, oidx)
    fin = np.isfinite(om)
    nanmatch = np.array_equal(np.isnan(mean), np.isnan(om))
    dm = np.abs(mean[fin] - om[fin]); dv = np.abs(var[fin] - ov[fin])
    rm = (dm / np.maximum(np.abs(om[fin]), 1e-300)).max() if fin.any() else 0
    rv = (dv / np.maximum(np.abs(ov[fin]), 1e-300)).max() if fin.any() else 0
    print(f"{name:28s} T={len(mean):8d} nn_ok={ok_nn} idx_ok={ok_idx} nan_ok={nanmatch} max|dmean|={dm.max() if fin.any() else 0:.3e} "
          f"relmean={rm:.3e} max|dvar|={dv.max() if fin.any() else 0:.3e} relvar={rv:.3e} gpu={t1-t0:.3f}s cpu={t2-t1:.3f}s", flush=True)
    if idx is not None and not ok_idx:
        bad = np.flatnonzero((idx != oidx).any(axis=1))
        print("   first mismatches:", bad[:5], idx[bad[0]], oidx[bad[0]])
    return mean, var

if __name__ == "__main__":
    S = gskrige.synth
    which = sys.argv[1:] or ["ref", "c2s", "c3s", "c5s", "c1", "ball"]
    if "ref" in which:
        coords=[np.array([25.,50.,75.]), np.array([25.,75.,50.])]; vals=np.array([1.,0.,1.])
        sup=gskrige.default_support_py([1.0,1.0],35.0)
        for k,rad in [(3,None),(3,100.0),(0,None)]:
            spec=gskrige.ProblemSpec(coords=coords, values=vals, grid_dims=(100,100), grid_origin=(0.5,0.5), grid_spacing=(1.,1.), support=sup,
                vario_kind=0, vario_range=35.0, max_neighbors=k, ball_radius=(rad if rad else float('nan')))
            compare(f"ref2d k={k} r={rad}", spec, search=O.SEARCH_BRUTE)
    if "c2s" in which:
        compare("C2 scale .2", S.config_spec("C2", scale=0.2))
    if "c3s" in which:
        compare("C3a scale .15", S.config_spec("C3a", scale=0.15))
        compare("C3b scale .15", S.config_spec("C3b", scale=0.15))
    if "c5s" in which:
        compare("C5 scale .06", S.config_spec("C5", scale=0.06))
    if "ball" in which:
        compare("C2 ball", S.config_spec("C2", scale=0.2, ball_radius=18.0, min_neighbors=4))
    if "c1" in which:
        compare("C1", S.config_spec("C1"))
    if "c2" in which:
        compare("C2 full", S.config_spec("C2"))
