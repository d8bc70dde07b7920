"""Shared helpers for the test suite: the golden cases and the reference's own test problems."""
import importlib.util
from pathlib import Path

import numpy as np

_spec = importlib.util.spec_from_file_location("make_golden", Path(__file__).parent / "golden" / "make_golden.py")
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
GOLDEN = np.load(Path(__file__).parent / "golden" / "kriging_small.npz")
CASES = make_golden.CASES
build_case = make_golden.build


def ref_problem_2d(gsk, k=0, radius=None):
    """ref test/estimation/krig.jl:25-28 — z=[1,0,1] at (25,25),(50,75),(75,50) on
    CartesianGrid((100,100),(0.5,0.5),(1.0,1.0)), GaussianVariogram(range=35, nugget=0)."""
    coords = [np.array([25.0, 50.0, 75.0]), np.array([25.0, 75.0, 50.0])]
    vals = np.array([1.0, 0.0, 1.0])
    sup = gsk.default_support_py([1.0, 1.0], 35.0)
    return gsk.ProblemSpec(coords=coords, values=vals, grid_dims=(100, 100), grid_origin=(0.5, 0.5),
                           grid_spacing=(1.0, 1.0), support=sup, vario_kind=gsk.VARIO_GAUSSIAN, vario_range=35.0,
                           max_neighbors=k, ball_radius=float("nan") if radius is None else radius)


def ref_problem_1d(gsk, k=0, radius=None):
    """ref test/estimation/krig.jl:6-8 — 11 samples at x=0:10:100 onto CartesianGrid(100)."""
    coords = [np.arange(0.0, 101.0, 10.0)]
    vals = np.array([0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.4, 0.3, 0.2, 0.1, 0.0])
    sup = gsk.default_support_py([1.0], 35.0)
    return gsk.ProblemSpec(coords=coords, values=vals, grid_dims=(100,), support=sup, vario_kind=gsk.VARIO_GAUSSIAN,
                           vario_range=35.0, max_neighbors=k, ball_radius=float("nan") if radius is None else radius)
