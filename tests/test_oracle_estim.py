"""The oracle's IDW / LWR bodies (oracle/gsk_oracle.c) against an independent numpy statement of
ref src/estimation/idw.jl:118-140 and src/estimation/lwr.jl:119-145, and the traversal-order plumbing."""
import numpy as np
import pytest


def _numpy_idw(coords, vals, centers, k, expo):
    X = np.stack(coords, 1)
    out_mu, out_sig = [], []
    for c in centers:
        d = np.sqrt(((X - c) ** 2).sum(1))
        idx = np.lexsort((np.arange(len(d)), d))[: (k or len(d))]
        ds = d[idx]
        ws = 1.0 / ds ** expo if expo != 1 else 1.0 / ds
        sw = ws.sum()
        if np.isinf(sw):
            j = np.flatnonzero(ds == 0)[0]
            out_mu.append(vals[idx[j]]); out_sig.append(0.0)
        else:
            out_mu.append(((ws / sw) * vals[idx]).sum()); out_sig.append(ds.min())
    return np.array(out_mu), np.array(out_sig)


def _numpy_lwr(coords, vals, centers, k):
    X = np.stack(coords, 1)
    out_mu, out_sig = [], []
    for c in centers:
        d = np.sqrt(((X - c) ** 2).sum(1))
        idx = np.lexsort((np.arange(len(d)), d))[: (k or len(d))]
        ds = d[idx]
        W = np.diag(np.exp(-3.0 * (ds / ds.max()) ** 2))
        Xl = np.hstack([np.ones((len(idx), 1)), X[idx]])
        A = Xl.T @ W @ Xl
        theta = np.linalg.solve(A, Xl.T @ W @ vals[idx])
        x0 = np.concatenate([[1.0], c])
        out_mu.append(theta @ x0)
        out_sig.append(np.linalg.norm(W @ Xl @ np.linalg.solve(A, x0)))
    return np.array(out_mu), np.array(out_sig)


@pytest.mark.parametrize("dim,grid,n,k", [(2, (12, 9), 60, 6), (3, (6, 5, 4), 80, 9), (2, (8, 8), 25, 0), (1, (40,), 15, 4)])
def test_oracle_idw_lwr_match_numpy(gsk, oracle, dim, grid, n, k):
    rng = np.random.default_rng(7 * dim + n)
    coords = [rng.uniform(0.0, g, n) for g in grid]
    vals = rng.standard_normal(n)
    base = gsk.ProblemSpec(coords=coords, values=vals, grid_dims=grid, max_neighbors=k)
    centers = np.stack(base.target_centers(), 1)
    for expo in (1.0, 2.0, 2.5):
        spec = gsk.ProblemSpec(coords=coords, values=vals, grid_dims=grid, max_neighbors=k, solver=gsk.SOLVER_IDW, idw_exponent=expo)
        mu, sig = oracle.krige(spec)
        nmu, nsig = _numpy_idw(coords, vals, centers, k, expo)
        np.testing.assert_allclose(mu, nmu, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(sig, nsig, rtol=1e-14)
    spec = gsk.ProblemSpec(coords=coords, values=vals, grid_dims=grid, max_neighbors=k, solver=gsk.SOLVER_LWR)
    mu, sig = oracle.krige(spec)
    nmu, nsig = _numpy_lwr(coords, vals, centers, k)
    np.testing.assert_allclose(mu, nmu, rtol=1e-8, atol=1e-9)      # normal equations in raw coordinates: cond·eps
    np.testing.assert_allclose(sig, nsig, rtol=1e-8, atol=1e-9)


def test_oracle_idw_reference_problem(gsk, oracle):
    """ref test/estimation/idw.jl:2-9 (the problem) and :67-73 (the data-location check, here on the scalar field)"""
    coords = [np.array([25.0, 50.0, 75.0]), np.array([25.0, 75.0, 50.0])]
    vals = np.array([1.0, 0.0, 1.0])
    spec = gsk.ProblemSpec(coords=coords, values=vals, grid_dims=(100, 100), max_neighbors=3, solver=gsk.SOLVER_IDW)
    mu, sig, nn, idx = oracle.krige(spec, want_neighbors=True)
    Z = mu.reshape((100, 100), order="F")
    assert abs(Z[24, 24] - 1.0) < 5e-2 and abs(Z[49, 74] - 0.0) < 5e-2 and abs(Z[74, 49] - 1.0) < 5e-2
    assert np.all(nn == 3) and np.all((mu >= 0) & (mu <= 1)) and np.all(sig > 0)
    # a sample exactly on a centroid: zero distance → its value, distance 0 (idw.jl:127-130)
    spec0 = gsk.ProblemSpec(coords=[np.array([24.5, 50.0]), np.array([24.5, 75.0])], values=np.array([7.0, 1.0]),
                            grid_dims=(100, 100), max_neighbors=2, solver=gsk.SOLVER_IDW)
    mu0, sig0 = oracle.krige(spec0)
    assert mu0[24 + 24 * 100] == 7.0 and sig0[24 + 24 * 100] == 0.0


def test_oracle_target_order(gsk, oracle):
    import copy
    spec = gsk.synth.config_spec("C2", scale=0.03)
    T = spec.n_targets
    order = np.random.default_rng(0).permutation(T).astype(np.int64)
    mu, var, nn, idx = oracle.krige(spec, want_neighbors=True)
    sp = copy.copy(spec)
    sp.target_order = order
    pm, pv, pn, pi = oracle.krige(sp, want_neighbors=True)
    assert np.array_equal(pm, mu[order]) and np.array_equal(pv, var[order]) and np.array_equal(pi, idx[order])


def test_traverse_orders(gsk):
    g = gsk.CartesianGrid(5, 3)
    assert gsk.traverse(g, gsk.LinearPath()) is None
    mg = gsk.traverse(g, gsk.MultiGridPath())
    assert sorted(mg.tolist()) == list(range(15)) and mg[0] == 0
    assert mg.tolist()[:3] == [0, 4, 2]            # coarse levels first: (0,0) at step 8; (4,0) at step 4; (2,0) at step 2
    r1, r2 = gsk.traverse(g, gsk.RandomPath(seed=3)), gsk.traverse(g, gsk.RandomPath(seed=3))
    assert np.array_equal(r1, r2) and sorted(r1.tolist()) == list(range(15))
    with pytest.raises(gsk.UnsupportedOption):
        gsk.traverse(gsk.PointSet([(0.0, 0.0)]), gsk.MultiGridPath())
