"""Wall-clock timings of sequential Gaussian simulation through the C ABI (gsk_sgs_plan / gsk_sgs_sample), host buffers
in and out. Usage: python scripts/dev/sgs_timing.py  (prints one line per case)"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import gskrige as gsk  # noqa: E402


def coords_of(dims):
    axes = [(np.arange(d) + 0.5) for d in dims]
    mesh = np.meshgrid(*axes, indexing="ij")
    return [np.ascontiguousarray(np.transpose(m, tuple(reversed(range(len(dims))))).ravel()) for m in mesh]


def case(ctx, dims, k, path, nreals, ndata=100):
    n = int(np.prod(dims))
    rng = np.random.default_rng(0)
    cs = coords_of(dims)
    order = np.arange(n) if path == "linear" else rng.permutation(n)
    data = rng.choice(n, ndata, replace=False)
    isdata = np.zeros(n, dtype=bool)
    isdata[data] = True
    visit = order[~isdata[order]]
    rank = np.full(n, -1, dtype=np.int64)
    rank[visit] = np.arange(len(visit))
    vals = np.where(isdata, rng.standard_normal(n), 0.0)
    kw = dict(vario_kind=gsk.VARIO_SPHERICAL, vario_range=20.0, max_neighbors=k)
    ctx.sgs_plan(cs, rank, **kw)
    t0 = time.perf_counter()
    ctx.sgs_plan(cs, rank, **kw)
    t_plan = time.perf_counter() - t0
    tp = ctx.timing()
    line = (f"SGS {'x'.join(map(str, dims))} ({n} locations, {ndata} data), k={k}, {path} path: plan {t_plan * 1e3:.1f} ms "
            f"[bins+upload {tp['ms_plan']:.1f}, sort+search {tp['ms_search']:.1f}, weights {tp['ms_solve']:.1f}, "
            f"levels {tp['ms_total'] - tp['ms_plan'] - tp['ms_search'] - tp['ms_solve']:.1f}]")
    for nr in nreals:
        z = rng.standard_normal((nr, n))
        ctx.sgs_sample(z[:1], values=vals)
        t0 = time.perf_counter()
        out = ctx.sgs_sample(z, values=vals)
        dt = time.perf_counter() - t0
        assert np.isfinite(out).all()
        line += f"; {nr} realisation(s) {dt * 1e3:.1f} ms ({nr * n / dt:.3e} locations/s; kernels {ctx.timing()['ms_solve']:.2f} ms, {ctx.timing()['launches']} launches)"
    print(line, flush=True)


if __name__ == "__main__":
    ctx = gsk.Context(0)
    case(ctx, (1000, 1000), 10, "random", (1, 16))
    case(ctx, (1000, 1000), 10, "linear", (1,))
    case(ctx, (512, 512), 10, "random", (1, 148, 592))
    case(ctx, (512, 512), 32, "random", (1, 148))
    case(ctx, (100, 100, 50), 16, "random", (1, 148))
    case(ctx, (1000000,), 4, "linear", (1, 148))
    case(ctx, (1000000,), 4, "random", (1, 148))
    ctx.close()
