"""Pins the CPU oracle against every known-answer test the reference holds for the Kriging path
(SURVEY §8c): test/estimation/krig.jl:35-37 (global), :50-52 (kNN k=3), :70-72 (ball k=3 r=100,
through LinearIndices → column-major x-fastest), the smoke-only 1-D problems (:6-19) and the
UI checks of test/ui.jl:7-37 (on the host mirror)."""
import warnings

import numpy as np
import pytest

from _cases import GOLDEN, ref_problem_1d, ref_problem_2d


@pytest.mark.parametrize("k,radius", [(0, None), (3, None), (3, 100.0)])
def test_reference_known_answers_2d(gsk, oracle, k, radius):
    spec = ref_problem_2d(gsk, k, radius)
    mean, var = oracle.krige(spec, search=oracle.SEARCH_BRUTE)
    Z = mean.reshape((100, 100), order="F")           # asarray(sol, :z)
    lin = lambda i, j: (i - 1) + (j - 1) * 100       # LinearIndices(size(grid))[i, j], 1-based → 0-based
    for i, j, expected in GOLDEN["ref_checks"]:
        i, j = int(i), int(j)
        assert abs(Z[i - 1, j - 1] - expected) < 1e-3          # krig.jl:35-37, 50-52
        assert abs(mean[lin(i, j)] - expected) < 1e-3          # krig.jl:70-72
    assert np.all(np.isfinite(var)) and var.min() >= 0.0


@pytest.mark.parametrize("k,radius", [(0, None), (3, None), (3, 100.0)])
def test_reference_smoke_1d(gsk, oracle, k, radius):
    mean, var = oracle.krige(ref_problem_1d(gsk, k, radius))
    assert mean.shape == (100,) and np.all(np.isfinite(mean)) and np.all(np.isfinite(var))
    assert mean.max() < 0.6 and mean.min() > -0.1


def test_searcher_ui(gsk):
    """ref test/ui.jl:7-23"""
    domain = gsk.PointSet(np.random.default_rng(0).random((2, 3)))
    m = gsk.searcher_ui(domain, 2, gsk.Euclidean(), None)
    assert isinstance(m, gsk.KNearestSearch) and gsk.maxneighbors(m) == 2
    m = gsk.searcher_ui(domain, 2, None, gsk.MetricBall(1.0))
    assert isinstance(m, gsk.KBallSearch) and gsk.maxneighbors(m) == 2
    m = gsk.searcher_ui(domain, None, gsk.Euclidean(), None)
    assert isinstance(m, gsk.KNearestSearch) and gsk.maxneighbors(m) == 3
    with pytest.warns(UserWarning, match=r"Invalid maximum number of neighbors\. Adjusting to 3\.\.\."):
        m = gsk.searcher_ui(domain, 4, gsk.Euclidean(), None)
    assert isinstance(m, gsk.KNearestSearch) and gsk.maxneighbors(m) == 3
    with pytest.warns(UserWarning):
        assert gsk.maxneighbors(gsk.searcher_ui(domain, 0, gsk.Euclidean(), None)) == 3


def test_kriging_ui(gsk):
    """ref test/ui.jl:29-37 — precedence drifts > degree > mean > ordinary"""
    grid = gsk.CartesianGrid(10, 10)
    g = gsk.GaussianVariogram()
    assert isinstance(gsk.kriging_ui(grid, g, None, None, None), gsk.OrdinaryKriging)
    assert isinstance(gsk.kriging_ui(grid, g, 0.0, None, None), gsk.SimpleKriging)
    uk = gsk.kriging_ui(grid, g, None, 2, None)
    assert isinstance(uk, gsk.UniversalKriging) and uk.dim == 2 and uk.degree == 2
    assert isinstance(gsk.kriging_ui(grid, g, None, None, [lambda x: 1]), gsk.ExternalDriftKriging)
    assert isinstance(gsk.kriging_ui(grid, g, 1.0, 1, [lambda x: 1]), gsk.ExternalDriftKriging)
    assert isinstance(gsk.kriging_ui(grid, g, 1.0, 1, None), gsk.UniversalKriging)
