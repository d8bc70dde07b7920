"""CPU checks of the oracle's restatement of sequential Gaussian simulation (ref src/simulation/seq.jl:102-135 with the
estimator and marginal of src/simulation/sgs.jl:62-69) and of the host mirror that prepares the path for it.

Known answers: (1) on a line with an exponential variogram, a linear path and one neighbour, the sequential loop IS the
AR(1) process v_i = μ + ρ(v_{i-1} − μ) + √(s(1 − ρ²))·z_i with ρ = exp(−3Δ/r); (2) the weights of any location solve the
Simple Kriging system of its masked neighbours (numpy.linalg.solve); (3) data are honoured exactly, as the reference's
own test asserts (test/simulation/sgs.jl:18-21)."""
import numpy as np


def test_sgs_oracle_ar1_known_answer(gsk, oracle):
    n, dx, r, s, mu = 200, 0.5, 4.0, 1.7, 3.0
    x = (np.arange(n) + 0.5) * dx
    z = np.random.default_rng(0).standard_normal(n)
    out, nn, idx, lam, sig = oracle.sgs([x], np.arange(n), vario_kind=gsk.VARIO_EXPONENTIAL, vario_range=r, vario_sill=s,
                                        mean=mu, max_neighbors=1, z=z)
    rho = np.exp(-3.0 * dx / r)
    want = np.empty(n)
    want[0] = mu + np.sqrt(s) * z[0]
    for i in range(1, n):
        want[i] = mu + rho * (want[i - 1] - mu) + np.sqrt(s * (1.0 - rho * rho)) * z[i]
    np.testing.assert_allclose(out, want, rtol=1e-12, atol=1e-12)
    assert nn[0] == 0 and (nn[1:] == 1).all() and np.array_equal(idx[1:, 0], np.arange(n - 1))


def test_sgs_oracle_weights_solve_the_masked_system(gsk, oracle):
    rng = np.random.default_rng(1)
    n, k = 300, 7
    pts = rng.uniform(0.0, 20.0, (2, n))
    data = rng.choice(n, 25, replace=False)
    order = rng.permutation(n)
    isdata = np.zeros(n, dtype=bool)
    isdata[data] = True
    visit = order[~isdata[order]]
    rank = np.full(n, -1, dtype=np.int64)
    rank[visit] = np.arange(len(visit))
    values = np.where(isdata, rng.standard_normal(n), 0.0)
    z = rng.standard_normal(n)
    s, r, nug, mu = 2.0, 6.0, 0.2, 0.5
    out, nn, idx, lam, sig = oracle.sgs(list(pts), rank, vario_kind=gsk.VARIO_SPHERICAL, vario_range=r, vario_sill=s,
                                        vario_nugget=nug, mean=mu, max_neighbors=k, ball_radius=5.0, values=values, z=z)
    assert np.array_equal(out[isdata], values[isdata])

    def cov(h):
        t = h / r
        g = np.where(h < r, (s - nug) * (1.5 * t - 0.5 * t ** 3), s - nug) + np.where(h > 0, nug, 0.0)
        return s - g

    for i in visit[::13]:
        # the mask: data and locations of lower rank, within the ball, the k nearest by (distance, index)
        elig = np.flatnonzero(isdata | ((rank >= 0) & (rank < rank[i])))
        d = np.sqrt(((pts[:, elig] - pts[:, [i]]) ** 2).sum(0))
        keep = np.lexsort((elig, d))[:k]
        keep = keep[d[keep] <= 5.0]
        nb = elig[keep]
        assert nn[i] == len(nb) and np.array_equal(idx[i, :len(nb)], nb)
        if len(nb) == 0:
            assert out[i] == mu + np.sqrt(s) * z[i]
            continue
        C = cov(np.sqrt(((pts[:, nb][:, :, None] - pts[:, nb][:, None, :]) ** 2).sum(0)))
        b = cov(d[keep])
        w = np.linalg.solve(C, b)
        np.testing.assert_allclose(lam[i, :len(nb)], w, rtol=1e-9, atol=1e-11)
        s2 = max(s - b @ w, 0.0)
        np.testing.assert_allclose(sig[i], np.sqrt(s2), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(out[i], mu + w @ (out[nb] - mu) + np.sqrt(s2) * z[i], rtol=1e-10, atol=1e-11)


class _PlanRecorder:
    def sgs_plan(self, coords, rank, **kw):
        self.coords, self.rank, self.kw = coords, np.asarray(rank), kw

    def sgs_sample(self, z, values=None):
        self.z, self.values = z, values
        return np.where(self.rank[None, :] < 0, values[None, :], z)


def test_sgs_host_mirror_prepares_path_and_data(gsk):
    """sgs.jl:56-89 / seq.jl:76-103: NearestInit puts the data into their cells, the path skips them, the searcher
    follows searcher_ui, and solve() returns `nreals` tables that carry the data values"""
    S = gsk.georef({"z": [1.0, 0.0, 1.0]}, np.array([[25.0, 50.0, 75.0], [25.0, 75.0, 50.0]]))
    D = gsk.CartesianGrid((100, 100), (0.5, 0.5), (1.0, 1.0))
    solver = gsk.SGS(z=dict(variogram=gsk.SphericalVariogram(range=35.0), neighborhood=gsk.MetricBall(30.0),
                            path=gsk.RandomPath(4)), rng=2017)
    rec = _PlanRecorder()
    sol = gsk.solve(gsk.SimulationProblem(S, D, "z", 3), solver, rec)
    assert len(sol) == 3
    cells = [24 * 100 + 24, 74 * 100 + 49, 49 * 100 + 74]
    assert (rec.rank[cells] == -1).all() and (rec.rank < 0).sum() == 3
    visit = gsk.traverse(D, gsk.RandomPath(4))
    visit = visit[~np.isin(visit, cells)]
    assert np.array_equal(np.argsort(rec.rank)[3:], visit)            # rank = position in the path among the others
    assert rec.kw["max_neighbors"] == 10 and rec.kw["ball_radius"] == 30.0 and rec.kw["min_neighbors"] == 1
    assert rec.kw["vario_kind"] == gsk.VARIO_SPHERICAL and rec.kw["mean"] == 0.0
    assert rec.z.shape == (3, 10000) and (rec.z[:, cells] == 0.0).all()
    for t in sol:
        col = np.asarray(t["z"])
        assert col[cells[0]] == 1.0 and col[cells[1]] == 0.0 and col[cells[2]] == 1.0
    # unconditional, default LinearPath, KNearestSearch
    rec2 = _PlanRecorder()
    gsk.solve(gsk.SimulationProblem(D, "z", 1), gsk.SGS(z=dict(maxneighbors=5), rng=1), rec2)
    assert np.array_equal(rec2.rank, np.arange(10000)) and np.isnan(rec2.kw["ball_radius"]) and rec2.kw["max_neighbors"] == 5
    assert rec2.kw["vario_kind"] == gsk.VARIO_GAUSSIAN
