"""GPU parity of sequential Gaussian simulation (ref src/simulation/sgs.jl, src/simulation/seq.jl) against the oracle's
restatement of the sequential loop: the masked neighbour sets must be identical, the Simple Kriging weights and
conditional standard deviations agree within the conditioning of the k×k systems, and realisations driven by the same
draws agree location by location. The reference's own test (test/simulation/sgs.jl) is mirrored at its full size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _grid_coords(dims, spacing=1.0):
    axes = [(np.arange(d) + 0.5) * spacing for d in dims]
    mesh = np.meshgrid(*axes, indexing="ij")
    # x fastest (column-major linear index, as the domain's elements are numbered)
    return [np.ascontiguousarray(np.transpose(m, tuple(reversed(range(len(dims))))).ravel()) for m in mesh]


def _rank(n, data_idx, order):
    rank = np.full(n, -1, dtype=np.int64)
    isdata = np.zeros(n, dtype=bool)
    isdata[data_idx] = True
    visit = order[~isdata[order]]
    rank[visit] = np.arange(len(visit))
    return rank


def _paths(gsk, dims, seed):
    dom = gsk.CartesianGrid(*dims) if len(dims) > 1 else gsk.CartesianGrid(dims[0])
    n = int(np.prod(dims))
    return {"linear": np.arange(n, dtype=np.int64), "random": gsk.traverse(dom, gsk.RandomPath(seed)),
            "multigrid": gsk.traverse(dom, gsk.MultiGridPath())}


CASES = [
    # dims, variogram kind, range, nugget, k, min_neighbors, ball radius, path, number of data
    ((40, 30), "spherical", 12.0, 0.0, 10, 1, None, "linear", 12),
    ((40, 30), "spherical", 12.0, 0.1, 10, 1, 6.0, "random", 12),
    ((33, 31), "exponential", 9.0, 0.0, 6, 3, 4.0, "multigrid", 0),
    ((32, 32), "gaussian", 3.0, 0.05, 8, 1, None, "random", 20),
    ((400,), "exponential", 15.0, 0.0, 4, 1, None, "random", 5),
    ((400,), "spherical", 25.0, 0.0, 5, 2, 10.0, "linear", 0),
    ((12, 11, 10), "spherical", 5.0, 0.0, 12, 1, None, "random", 30),
    ((12, 11, 10), "exponential", 4.0, 0.2, 26, 1, None, "multigrid", 10),   # heap top-k in the search
    ((24, 20), "exponential", 8.0, 0.1, 40, 1, None, "random", 15),          # two neighbours per lane, global scratch
    ((16, 16), "spherical", 6.0, 0.0, 64, 1, None, "random", 8),             # the largest k
]


@pytest.mark.parametrize("dims,vk,rng_,nug,k,minn,radius,path,ndata", CASES)
def test_sgs_vs_oracle(gsk, ctx, oracle, dims, vk, rng_, nug, k, minn, radius, path, ndata):
    rng = np.random.default_rng(hash((dims, k, ndata)) % (2 ** 32))
    n = int(np.prod(dims))
    coords = _grid_coords(dims)
    kind = {"gaussian": gsk.VARIO_GAUSSIAN, "spherical": gsk.VARIO_SPHERICAL, "exponential": gsk.VARIO_EXPONENTIAL}[vk]
    data_idx = rng.choice(n, ndata, replace=False) if ndata else np.zeros(0, dtype=np.int64)
    rank = _rank(n, data_idx, _paths(gsk, dims, 7)[path])
    values = np.zeros(n)
    values[data_idx] = 2.0 + rng.standard_normal(ndata)
    z = rng.standard_normal((3, n))
    kw = dict(vario_kind=kind, vario_range=rng_, vario_sill=1.5, vario_nugget=nug, mean=2.0, min_neighbors=minn,
              max_neighbors=k, ball_radius=float("nan") if radius is None else radius)
    ctx.sgs_plan(coords, rank, **kw)
    nn, idx, lam, sig = ctx.sgs_weights()
    reals = ctx.sgs_sample(z, values=values)
    assert ctx.timing()["launches"] >= 1
    # the conditioning of the k×k covariance systems bounds how closely two factorisations can agree
    wtol = 1e-6 if vk == "gaussian" else 1e-8
    for r in range(z.shape[0]):
        out, onn, oidx, olam, osig = oracle.sgs(coords, rank, values=values, z=z[r], **kw)
        if r == 0:
            sim = rank >= 0
            assert np.array_equal(nn[sim], onn[sim])
            assert np.array_equal(idx[sim], oidx[sim])
            np.testing.assert_allclose(lam[sim], olam[sim], rtol=wtol, atol=wtol)
            np.testing.assert_allclose(sig[sim], osig[sim], rtol=wtol, atol=wtol)
            if minn > 1 or radius is not None:
                assert (onn[sim] == 0).any()          # the marginal branch (seq.jl:108-110) is exercised
        assert np.array_equal(reals[r][~sim], values[~sim])
        np.testing.assert_allclose(reals[r], out, rtol=100 * wtol, atol=100 * wtol)


def test_sgs_reference_testset(gsk, ctx, oracle):
    """ref test/simulation/sgs.jl:1-22 — 100×100 grid, three data, SphericalVariogram(range=35), MetricBall(30):
    conditional and unconditional problems through solve(); every realisation honours the data exactly."""
    S = gsk.georef({"z": [1.0, 0.0, 1.0]}, np.array([[25.0, 50.0, 75.0], [25.0, 75.0, 50.0]]))
    D = gsk.CartesianGrid((100, 100), (0.5, 0.5), (1.0, 1.0))
    N = 3
    solver = gsk.SGS(z=dict(variogram=gsk.SphericalVariogram(range=35.0), neighborhood=gsk.MetricBall(30.0)), rng=2017)
    sol1 = gsk.solve(gsk.SimulationProblem(S, D, "z", N), solver, ctx)
    sol2 = gsk.solve(gsk.SimulationProblem(D, "z", N), solver, ctx)
    assert len(sol1) == N and len(sol2) == N
    lin = lambda i, j: (j - 1) * 100 + (i - 1)                       # LinearIndices(size(D))[i, j], 0-based
    for r in range(N):
        z1 = np.asarray(sol1[r]["z"])
        assert z1[lin(25, 25)] == 1.0 and z1[lin(50, 75)] == 0.0 and z1[lin(75, 50)] == 1.0
        assert np.isfinite(z1).all() and np.isfinite(np.asarray(sol2[r]["z"])).all()
    # the same problem, the same draws, through the oracle
    from gskrige import simulation as sim
    pre = sim.preprocess_sgs(gsk.SimulationProblem(S, D, "z", 1), solver, "z", ctx)
    zz = np.random.default_rng(5).standard_normal(pre["npts"])
    got = ctx.sgs_sample(zz, values=pre["values"])
    out, *_ = oracle.sgs(D.centroids(), pre["rank"], vario_kind=gsk.VARIO_SPHERICAL, vario_range=35.0, mean=0.0,
                         min_neighbors=1, max_neighbors=10, ball_radius=30.0, values=pre["values"], z=zz)
    np.testing.assert_allclose(got, out, rtol=1e-6, atol=1e-6)


def test_sgs_ensemble_and_statistics(gsk, ctx):
    """64 realisations in one call equal 64 single calls; an unconditional ensemble reproduces mean and sill"""
    dims = (64, 64)
    n = dims[0] * dims[1]
    coords = _grid_coords(dims)
    rank = _rank(n, np.zeros(0, dtype=np.int64), np.random.default_rng(3).permutation(n))
    ctx.sgs_plan(coords, rank, vario_kind=gsk.VARIO_EXPONENTIAL, vario_range=10.0, vario_sill=2.0, mean=-1.0,
                 max_neighbors=16)
    z = np.random.default_rng(11).standard_normal((64, n))
    ens = ctx.sgs_sample(z)
    for r in (0, 17, 63):
        assert np.array_equal(ctx.sgs_sample(z[r]), ens[r])
    assert abs(ens.mean() + 1.0) < 0.15
    assert abs(ens.var() - 2.0) < 0.3
    # neighbouring cells are correlated as the variogram says: γ(1) = sill·(1 − exp(−3/10))
    f = ens.reshape(64, dims[1], dims[0])
    g1 = 0.5 * np.mean((f[:, :, 1:] - f[:, :, :-1]) ** 2)
    assert abs(g1 - 2.0 * (1.0 - np.exp(-0.3))) < 0.08


def test_sgs_errors(gsk, ctx):
    coords = _grid_coords((10, 10))
    good = np.arange(100, dtype=np.int64)
    kw = dict(vario_kind=gsk.VARIO_SPHERICAL, vario_range=5.0)
    gsk.Context(0).close()
    fresh = gsk.Context(0)
    with pytest.raises(gsk.GskError):
        fresh.lib.gsk_sgs_sample  # noqa: B018 (symbol exists)
        fresh._sgs_n = 100
        fresh.sgs_sample(np.zeros(100))                               # before any plan
    dup = good.copy()
    dup[1] = 0
    with pytest.raises(gsk.GskError):
        fresh.sgs_plan(coords, dup, **kw)
    with pytest.raises(gsk.GskError):
        fresh.sgs_plan(coords, np.full(100, -1), **kw)                # nothing to simulate
    with pytest.raises(gsk.GskError):
        fresh.sgs_plan(coords, good, max_neighbors=65, **kw)
    # a Kriging plan made afterwards on the same context still works (the buffers are shared)
    fresh.sgs_plan(coords, good, **kw)
    spec = gsk.synth.config_spec("C2", scale=0.05)
    m, v = fresh.krige(spec)
    assert np.isfinite(m).all()
    with pytest.raises(gsk.GskError):
        fresh.sgs_sample(np.zeros(100))                               # the SGS plan was replaced
    fresh.close()


def test_sgs_ar1_known_answer(gsk, ctx):
    """on a line with an exponential variogram, a linear path and one neighbour the loop is the AR(1) process
    v_i = μ + ρ(v_{i-1} − μ) + √(s(1 − ρ²))·z_i, ρ = exp(−3Δ/r)"""
    n, dx, r, s, mu = 5000, 0.5, 4.0, 1.7, 3.0
    x = (np.arange(n) + 0.5) * dx
    z = np.random.default_rng(0).standard_normal(n)
    ctx.sgs_plan([x], np.arange(n), vario_kind=gsk.VARIO_EXPONENTIAL, vario_range=r, vario_sill=s, mean=mu, max_neighbors=1)
    out = ctx.sgs_sample(z)
    rho = np.exp(-3.0 * dx / r)
    want = np.empty(n)
    want[0] = mu + np.sqrt(s) * z[0]
    for i in range(1, n):
        want[i] = mu + rho * (want[i - 1] - mu) + np.sqrt(s * (1.0 - rho * rho)) * z[i]
    np.testing.assert_allclose(out, want, rtol=1e-10, atol=1e-10)


def test_sgs_device_buffers(gsk, ctx):
    """gsk_sgs_sample_device on the caller's device buffers gives the bytes of the host-buffer call; solve() with
    device_draws honours the data"""
    import torch
    dims = (50, 40)
    n = dims[0] * dims[1]
    coords = _grid_coords(dims)
    rng = np.random.default_rng(2)
    data_idx = rng.choice(n, 9, replace=False)
    rank = _rank(n, data_idx, rng.permutation(n))
    values = np.zeros(n)
    values[data_idx] = rng.standard_normal(9)
    z = rng.standard_normal((5, n))
    ctx.sgs_plan(coords, rank, vario_kind=gsk.VARIO_EXPONENTIAL, vario_range=9.0, max_neighbors=12)
    want = ctx.sgs_sample(z, values=values)
    dz = torch.from_numpy(z).cuda()
    dv = torch.from_numpy(values).cuda()
    dout = torch.empty_like(dz)
    torch.cuda.synchronize()
    ctx.sgs_sample_device(5, dv.data_ptr(), dz.data_ptr(), dout.data_ptr())
    ctx.synchronize()
    assert np.array_equal(dout.cpu().numpy(), want)
    with pytest.raises(gsk.GskError):
        ctx.sgs_sample_device(5, 0, dz.data_ptr(), dout.data_ptr())   # the plan has data: values are required
    from gskrige import simulation as sim
    S = gsk.georef({"z": [1.0, 0.0, 1.0]}, np.array([[25.0, 50.0, 75.0], [25.0, 75.0, 50.0]]))
    D = gsk.CartesianGrid((100, 100), (0.5, 0.5), (1.0, 1.0))
    solver = gsk.SGS(z=dict(variogram=gsk.SphericalVariogram(range=35.0), path=gsk.RandomPath(1)), rng=3)
    sol = sim.solve_sgs(gsk.SimulationProblem(S, D, "z", 4), solver, ctx, device_draws=True)
    for t in sol:
        col = np.asarray(t["z"])
        assert col[24 * 100 + 24] == 1.0 and col[74 * 100 + 49] == 0.0 and col[49 * 100 + 74] == 1.0
        assert np.isfinite(col).all() and 0.3 < col.std() < 3.0
