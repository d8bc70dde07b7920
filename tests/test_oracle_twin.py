"""Cross-checks the C oracle (hand-written LU / Cholesky) against the committed golden vectors, which
come from the independent numpy restatement that calls LAPACK dsytrf (Bunch–Kaufman) / dpotrf — the
factorisations Julia uses for the reference. Also: KD-tree search == brute force."""
import numpy as np
import pytest

from _cases import CASES, GOLDEN, build_case
from conftest import assert_parity


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden(gsk, oracle, case):
    name = case[0]
    spec = build_case(case)
    mean, var, nn, idx = oracle.krige(spec, search=oracle.SEARCH_BRUTE, want_neighbors=True)
    assert np.array_equal(nn, GOLDEN[f"{name}/nn"])
    if spec.params["max_neighbors"] > 0:
        assert np.array_equal(idx, GOLDEN[f"{name}/idx"])
    # Gaussian systems are ill-conditioned (cond ~1e7 with the 1e-6 nugget): LU vs Bunch–Kaufman
    # legitimately differ by ~cond·eps; the other families agree to rtol 1e-9.
    gauss = spec.params["vario_kind"] == gsk.VARIO_GAUSSIAN
    assert_parity(mean, var, GOLDEN[f"{name}/mean"], GOLDEN[f"{name}/var"], scale=np.abs(spec.values).max(),
                  atol_mean=2e-8 if gauss else None, atol_var=2e-9 if gauss else None)


@pytest.mark.parametrize("name,scale", [("C2", 0.08), ("C3a", 0.08), ("C5", 0.04)])
def test_kdtree_equals_brute(gsk, oracle, name, scale):
    spec = gsk.synth.config_spec(name, scale=scale)
    a = oracle.search(spec, search=oracle.SEARCH_BRUTE)
    b = oracle.search(spec, search=oracle.SEARCH_KDTREE)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    # sorted ascending by (d², idx)
    d2 = a[2]
    assert np.all(np.diff(d2, axis=1) >= 0)


def test_ties_fall_to_lower_index(gsk, oracle):
    """Four samples at the corners of a square, target at its centre: all distances tie."""
    coords = [np.array([0.0, 1.0, 0.0, 1.0, 5.0]), np.array([0.0, 0.0, 1.0, 1.0, 5.0])]
    spec = gsk.ProblemSpec(coords=coords, values=np.arange(5.0), points=[np.array([0.5]), np.array([0.5])],
                           vario_kind=gsk.VARIO_SPHERICAL, vario_range=3.0, max_neighbors=2)
    for kind in (oracle.SEARCH_BRUTE, oracle.SEARCH_KDTREE):
        nn, idx, d2 = oracle.search(spec, search=kind)
        assert idx.tolist() == [[0, 1]] and nn.tolist() == [2]


def test_uk_exponents_order(gsk, oracle):
    """GeoStatsModels UKexps: degree 1 → x, y, (z), 1; degree 2 (2-D) → x², y², x, y, xy, 1"""
    import numpy_twin as TW
    assert oracle.uk_exponents(1, 2).tolist() == [[1, 0], [0, 1], [0, 0]]
    assert oracle.uk_exponents(1, 3).tolist() == [[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0]]
    assert oracle.uk_exponents(2, 2).tolist() == [[2, 0], [0, 2], [1, 0], [0, 1], [1, 1], [0, 0]]
    for deg in (0, 1, 2):
        for dim in (1, 2, 3):
            assert oracle.uk_exponents(deg, dim).tolist() == TW.uk_exponents(deg, dim).tolist()
            assert gsk.uk_exponents(deg, dim).tolist() == TW.uk_exponents(deg, dim).tolist()
