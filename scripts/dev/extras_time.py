"""Throughput of the §8(f) rows on one B200 (documentation figures, not bench values): IDW / LWR on the C2 shape,
FFTGS conditional realisations on 512x512 with 1000 data, LUGS on 100x100 with 50 data."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import gskrige as gs
ctx = gs.Context(0)
base = gs.synth.config_spec("C2")
for name, sv, k in (("IDW k=20", gs.SOLVER_IDW, 20), ("LWR k=20", gs.SOLVER_LWR, 20), ("IDW all 10000 samples", gs.SOLVER_IDW, 0)):
    spec = gs.ProblemSpec(coords=base.coords, values=base.values, grid_dims=base.grid_dims, solver=sv, max_neighbors=k)
    T = spec.n_targets if k else 100_000
    sl = spec.with_slab(0, T)
    d_m = torch.empty(T, dtype=torch.float64, device="cuda"); d_v = torch.empty_like(d_m)
    ctx.plan(sl)
    for _ in range(3):
        ctx.execute(0, T, d_m.data_ptr(), d_v.data_ptr())
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.execute(0, T, d_m.data_ptr(), d_v.data_ptr())
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {T / dt:.3e} locations/s ({dt * 1e3:.3f} ms per {T} targets)", flush=True)
rng = np.random.default_rng(1)
grid = gs.CartesianGrid(512, 512)
nd = 1000
data = gs.georef({"z": rng.standard_normal(nd)}, np.stack([rng.uniform(0, 512, nd), rng.uniform(0, 512, nd)], 0))
prob = gs.SimulationProblem(data, grid, "z", 1)
solver = gs.FFTGS(z=dict(variogram=gs.SphericalVariogram(range=40.0), maxneighbors=20), rng=2)
t0 = time.perf_counter(); pre = gs.simulation.preprocess_fftgs(prob, solver, ctx); t1 = time.perf_counter()
ts = []
for _ in range(6):
    a = time.perf_counter(); gs.simulation.solvesingle_fftgs(prob, solver, pre, ctx); ts.append(time.perf_counter() - a)
print(f"FFTGS 512x512, 1000 data, SK maxneighbors=20: preprocess {1e3 * (t1 - t0):.1f} ms, realisations {[round(1e3 * t, 1) for t in ts]} ms "
      f"(first re-plans on the data cells; the rest are values-only updates with resident neighbour lists)", flush=True)
grid = gs.CartesianGrid(100, 100)
data = gs.georef({"z": rng.standard_normal(50)}, np.stack([rng.uniform(0, 100, 50), rng.uniform(0, 100, 50)], 0))
prob = gs.SimulationProblem(data, grid, "z", 1)
lug = gs.LUGS(z=dict(variogram=gs.SphericalVariogram(range=20.0)), rng=3)
t0 = time.perf_counter(); pre = gs.simulation.preprocess_lugs(prob, lug, "z", ctx); t1 = time.perf_counter()
ts = []
for _ in range(5):
    a = time.perf_counter(); gs.simulation.lusim(ctx, pre, rng.standard_normal(len(pre["slocs"]))); ts.append(time.perf_counter() - a)
print(f"LUGS 100x100 (10 000 points, 50 data): factorisation {1e3 * (t1 - t0):.1f} ms, realisations {[round(1e3 * t, 2) for t in ts]} ms", flush=True)
