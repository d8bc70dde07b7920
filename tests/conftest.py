import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gsk():
    """The host package (import name gskrige) with libgskrige.so built in-tree."""
    import gskrige
    if not gskrige.LIB_PATH.exists():
        from gskrige import build as _b
        _b.build()
    return gskrige


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def ctx(gsk):
    """One CUDA context for the GPU tests; fails loudly if the extension or the device is missing."""
    return gsk.Context(0)


# tolerances of the parity tests (north_star: rtol 1e-9 in Float64). The absolute floors cover
# outputs that pass through zero (means) and the cancellation in sill − b·w (variances).
RTOL = 1e-9


def assert_parity(mean, var, omean, ovar, *, scale=1.0, sill=1.0, atol_mean=None, atol_var=None):
    import numpy as np
    am = 1e-12 * scale if atol_mean is None else atol_mean
    av = 1e-12 * sill if atol_var is None else atol_var
    assert np.array_equal(np.isnan(mean), np.isnan(omean))
    assert np.array_equal(np.isnan(var), np.isnan(ovar))
    ok = ~np.isnan(omean)
    np.testing.assert_allclose(mean[ok], omean[ok], rtol=RTOL, atol=am)
    np.testing.assert_allclose(var[ok], ovar[ok], rtol=RTOL, atol=av)
