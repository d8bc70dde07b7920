"""One SGS plan + one 16-realisation sample on a 1000x1000 grid (random path, k = 10): the command the ncu launch list
of profiles/r02_sgs_launches.txt was taken from."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import gskrige as gsk  # noqa: E402

dims = (1000, 1000)
n = dims[0] * dims[1]
rng = np.random.default_rng(0)
gx, gy = np.meshgrid(np.arange(dims[0]) + 0.5, np.arange(dims[1]) + 0.5)
cs = [gx.ravel().copy(), gy.ravel().copy()]
data = rng.choice(n, 100, replace=False)
isdata = np.zeros(n, dtype=bool)
isdata[data] = True
order = rng.permutation(n)
visit = order[~isdata[order]]
rank = np.full(n, -1, dtype=np.int64)
rank[visit] = np.arange(len(visit))
vals = np.where(isdata, rng.standard_normal(n), 0.0)
ctx = gsk.Context(0)
ctx.sgs_plan(cs, rank, vario_kind=gsk.VARIO_SPHERICAL, vario_range=20.0, max_neighbors=10)
out = ctx.sgs_sample(rng.standard_normal((16, n)), values=vals)
print("ok", float(out.std()), ctx.timing())
ctx.close()
