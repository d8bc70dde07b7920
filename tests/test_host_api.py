"""CPU-side checks of the boundary: libgskrige.so loads and exports every symbol include/gskrige.h
declares, the ctypes struct matches the C layout, the host helpers agree with their Python
statements, and the host mirror reproduces the reference's host logic (krig.jl:76-164) — all
without a compute call (there is no GPU here and no CPU fallback to call)."""
import ctypes
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(gsk):
    lib = gsk.load_library()
    header = (ROOT / "include" / "gskrige.h").read_text()
    declared = re.findall(r"GSK_API\s+[\w\s\*]+?\b(gsk_\w+)\s*\(", header)
    assert sorted(set(declared)) == sorted(gsk.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gsk_abi_version() == 2


def test_struct_layout_matches_c(gsk, tmp_path):
    from gskrige._abi import GskProblem, GskTiming
    fields = [f for f, _ in GskProblem._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/gskrige.h"', 'int main(void){',
             'printf("%zu\\n", sizeof(gsk_problem));', 'printf("%zu\\n", sizeof(gsk_timing));']
    lines += [f'printf("%zu\\n", offsetof(gsk_problem, {f}));' for f in fields]
    lines += ['return 0;}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == ctypes.sizeof(GskProblem)
    assert out[1] == ctypes.sizeof(GskTiming)
    assert out[2:] == [getattr(GskProblem, f).offset for f in fields]


def test_compute_fails_loudly_without_gpu(gsk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gsk.GskError, match="no CUDA device|CUDA"):
        gsk.Context(0)


def test_default_support_helper(gsk):
    for spacing, rng in [([1.0], 35.0), ([1.0, 1.0], 35.0), ([1.0, 1.0, 1.0], 30.0), ([2.0, 1.0], 0.9), ([0.5, 0.5], 50.0)]:
        a = gsk.default_support(spacing, rng)
        b = gsk.default_support_py(spacing, rng)
        assert len(a) == len(b) == len(spacing)
        for x, y in zip(a, b):
            np.testing.assert_allclose(x, y, rtol=0, atol=1e-15)
    s = gsk.default_support([1.0, 1.0], 35.0)
    assert len(s[0]) == 9 and sorted(set(np.round(s[0], 12))) == [-0.25, 0.0, 0.25]   # 3 per axis at ±¼ side


def test_synthetic_generator_is_deterministic(gsk):
    u = gsk.synth.uniform01(42, 5)
    assert np.all((u >= 0) & (u < 1))
    # SplitMix64 known answers for seed 42 … first output
    z = gsk.synth.splitmix64(0, 2)
    assert int(z[0]) == 0xE220A8397B1DCDAF and int(z[1]) == 0x6E789E6AA1B965F4
    a = gsk.synth.config_spec("C2", scale=0.05)
    b = gsk.synth.config_spec("C2", scale=0.05)
    assert np.array_equal(a.coords[0], b.coords[0]) and np.array_equal(a.values, b.values)
    assert abs(gsk.synth.algorithmic_flops_per_target(gsk.synth.config_spec("C2", scale=0.05)) - 13060) < 1


class _FakeCtx:
    """Stands in for the CUDA context to observe what the host layer hands across the ABI."""

    def __init__(self):
        self.specs = []

    def krige(self, spec, want_neighbors=False, want_nneigh=False):
        self.specs.append(spec)
        self.asked_lists = want_neighbors
        _, count = spec.slab
        mean, var = np.zeros(count), np.ones(count)
        if want_neighbors or want_nneigh:
            nn = np.full(count, spec.params["max_neighbors"] or spec.n_samples, dtype=np.int32)
            nn[:2] = 0
            return (mean, var, nn, None) if want_neighbors else (mean, var, nn)
        return mean, var


def test_host_mirror_solve_flow(gsk):
    """krig.jl:76-164: missing filtering, unit adjustment, estimator/searcher choice, result columns."""
    table = {"z": gsk.Quantities([1.0, None, 0.0, 1.0], gsk.degC)}
    data = gsk.georef(table, [(25.0, 25.0), (1.0, 1.0), (50.0, 75.0), (75.0, 50.0)])
    grid = gsk.CartesianGrid((10, 10), (0.5, 0.5), (1.0, 1.0))
    problem = gsk.EstimationProblem(data, grid, "z")
    solver = gsk.KrigingSolver(z=dict(variogram=gsk.GaussianVariogram(range=35.0), maxneighbors=3,
                                      neighborhood=gsk.MetricBall(100.0), minneighbors=2))
    fake = _FakeCtx()
    sol = gsk.solve(problem, solver, ctx=fake)
    spec = fake.specs[0]
    assert spec.n_samples == 3                                   # the missing sample is dropped (krig.jl:97-107)
    np.testing.assert_allclose(spec.values, np.array([1.0, 0.0, 1.0]) + 273.15)   # °C → K (utils.jl:10-15)
    assert spec.params["max_neighbors"] == 3 and spec.params["ball_radius"] == 100.0
    assert spec.params["estimator"] == gsk.EST_ORDINARY and spec.params["min_neighbors"] == 2
    assert spec.support[0].shape == (9,)                         # block support of a 2-D cell
    assert sol.names() == ["z", "z_variance"]                    # krig.jl:159-163
    assert sol.z.unit == gsk.K and repr(sol["z_variance"].unit) == "K^2"
    assert np.ma.getmaskarray(sol.z.values)[:2].all() and not np.ma.getmaskarray(sol.z.values)[2:].any()
    assert gsk.asarray(sol, "z").shape == (10, 10)
    assert fake.asked_lists is False                             # only the neighbour COUNTS cross back, never the lists
    # a non-linear path: the visiting order crosses the ABI (the reference returns its results in that order)
    fake = _FakeCtx()
    gsk.solve(problem, gsk.KrigingSolver(z=dict(variogram=gsk.GaussianVariogram(range=35.0), maxneighbors=3,
                                                path=gsk.MultiGridPath())), ctx=fake)
    order = fake.specs[0].target_order
    assert order is not None and sorted(order.tolist()) == list(range(100)) and order[0] == 0
    assert np.array_equal(order, gsk.traverse(grid, gsk.MultiGridPath()))
    fake = _FakeCtx()
    gsk.solve(problem, gsk.KrigingSolver(z=dict(maxneighbors=3, path=gsk.MultiGridPath())), ctx=fake, path_order=False)
    assert fake.specs[0].target_order is None


def test_host_mirror_idw_lwr(gsk):
    """idw.jl:59-148 / lwr.jl:62-152: parameters, neighbour clamp, unit handling, column names."""
    table = {"T": gsk.Quantities([-272.15, None, -273.15, -272.15], gsk.degC)}
    data = gsk.georef(table, [(25.0, 25.0), (1.0, 1.0), (50.0, 75.0), (75.0, 50.0)])
    grid = gsk.CartesianGrid(5, 5)
    problem = gsk.EstimationProblem(data, grid, "T")
    fake = _FakeCtx()
    sol = gsk.solve(problem, gsk.IDWSolver(), ctx=fake)             # idw.jl:36-41: maxneighbors === nothing → every sample
    spec = fake.specs[0]
    assert spec.params["solver"] == gsk.SOLVER_IDW and spec.params["max_neighbors"] == 0 and spec.params["idw_exponent"] == 1.0
    assert spec.n_samples == 3
    np.testing.assert_allclose(spec.values, [1.0, 0.0, 1.0], atol=1e-12)   # °C → K
    assert sol.names() == ["T", "T_distance"] and gsk.elunit(sol["T"]) == gsk.K and gsk.elunit(sol["T_distance"]) == gsk.NoUnits
    fake = _FakeCtx()
    sol = gsk.solve(problem, gsk.LWRSolver(T=dict(maxneighbors=2, neighborhood=gsk.MetricBall(30.0))), ctx=fake)
    spec = fake.specs[0]
    assert spec.params["solver"] == gsk.SOLVER_LWR and spec.params["max_neighbors"] == 2 and spec.params["ball_radius"] == 30.0
    assert sol.names() == ["T", "T_variance"] and repr(gsk.elunit(sol["T_variance"])) == "K^2"
    with pytest.raises(AssertionError, match="exponent must be positive"):        # idw.jl:96
        gsk.solve(problem, gsk.IDWSolver(T=dict(exponent=0)), ctx=_FakeCtx())
    with pytest.raises(AssertionError, match="invalid min/max number of neighbors"):   # idw.jl:97
        gsk.solve(problem, gsk.IDWSolver(T=dict(minneighbors=3, maxneighbors=2)), ctx=_FakeCtx())
    with pytest.warns(UserWarning, match="Adjusting to 3"):
        f = _FakeCtx()
        gsk.solve(problem, gsk.LWRSolver(T=dict(maxneighbors=9)), ctx=f)
    assert f.specs[0].params["max_neighbors"] == 3
    with pytest.raises(TypeError):
        gsk.IDWSolver(T=dict(power=2))


def test_host_mirror_global_and_estimators(gsk):
    data = gsk.georef({"a": [1.0, 2.0, 3.0], "b": [0.0, 1.0, 0.5]}, np.array([[0.0, 1.0, 2.0], [0.0, 1.0, 0.0]]))
    pts = gsk.PointSet([(0.5, 0.5), (1.5, 0.5)])
    problem = gsk.EstimationProblem(data, pts, ("a", "b"))
    solver = gsk.KrigingSolver(("a", dict(variogram=gsk.SphericalVariogram(range=3.0), mean=2.0)),
                               ("b", dict(variogram=gsk.ExponentialVariogram(range=2.0), degree=1, maxneighbors=2)))
    fake = _FakeCtx()
    sol = gsk.solve(problem, solver, ctx=fake)
    a, b = fake.specs
    assert a.params["max_neighbors"] == 0 and a.params["estimator"] == gsk.EST_SIMPLE and a.params["sk_mean"] == 2.0
    assert b.params["estimator"] == gsk.EST_UNIVERSAL and b.params["uk_degree"] == 1 and b.params["max_neighbors"] == 2
    assert a.support[0].shape == (1,) and a.grid_dims is None     # PointSet targets → point support
    assert sol.names() == ["a", "b", "a_variance", "b_variance"]
    # default parameters for a variable that is not listed (krig.jl:53-58): Gaussian, global OK
    fake2 = _FakeCtx()
    gsk.solve(gsk.EstimationProblem(data, pts, "a"), gsk.KrigingSolver(), ctx=fake2)
    assert fake2.specs[0].params["vario_kind"] == gsk.VARIO_GAUSSIAN and fake2.specs[0].params["max_neighbors"] == 0


def test_host_mirror_errors(gsk):
    data = gsk.georef({"z": [None, None]}, [(0.0, 0.0), (1.0, 1.0)])
    grid = gsk.CartesianGrid(4, 4)
    with pytest.raises(AssertionError, match="all samples of z are missing, aborting..."):   # krig.jl:100-102
        gsk.solve(gsk.EstimationProblem(data, grid, "z"), gsk.KrigingSolver(), ctx=_FakeCtx())
    ok = gsk.georef({"z": [1.0, 2.0]}, [(0.0, 0.0), (1.0, 1.0)])
    prob = gsk.EstimationProblem(ok, grid, "z")
    with pytest.raises(gsk.UnsupportedOption):
        gsk.solve(prob, gsk.KrigingSolver(z=dict(drifts=[lambda x: 1.0])), ctx=_FakeCtx())
    with pytest.raises(gsk.UnsupportedOption):
        gsk.solve(prob, gsk.KrigingSolver(z=dict(maxneighbors=2, neighborhood=gsk.MetricBall((1.0, 2.0)))), ctx=_FakeCtx())
    with pytest.raises(gsk.UnsupportedOption):
        gsk.solve(prob, gsk.KrigingSolver(z=dict(degree=3)), ctx=_FakeCtx())
    with pytest.raises(TypeError):
        gsk.KrigingSolver(z=dict(varioogram=1))
    with pytest.warns(UserWarning, match="Adjusting to 2"):     # ui.jl:19, then the clamped k crosses the ABI
        f = _FakeCtx()
        gsk.solve(prob, gsk.KrigingSolver(z=dict(maxneighbors=5)), ctx=f)
    assert f.specs[0].params["max_neighbors"] == 2


def test_plain_c_host_compiles_against_the_header(gsk, tmp_path):
    """examples/krige_c.c is a C99 host (what a ccall-style binding does); it must build with the header alone."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    lib_dir = root / "geostatssolvers.jl_b200" / "csrc"
    exe = tmp_path / "krige_c"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{root / 'include'}",
                        str(root / "examples" / "krige_c.c"), f"-L{lib_dir}", "-lgskrige", f"-Wl,-rpath,{lib_dir}", "-lm",
                        "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)      # usage error, no compute call
    assert r.returncode == 2 and "usage" in r.stderr


def test_julia_shim_struct_matches_the_header(gsk):
    """julia/GSKrige.jl cannot run here (no julia binary); at least its `GskProblem` must list the fields of
    include/gskrige.h's gsk_problem in the same order with types of the same width (ccall passes it by reference)."""
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    hdr = (root / "include" / "gskrige.h").read_text()
    body = re.search(r"typedef struct gsk_problem \{(.*?)\} gsk_problem;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const double \*|const int64_t \*|double|int32_t|int64_t|uint32_t)\s*(.*)", decl)
        assert m, decl
        ctype = m.group(1).strip()
        for name in m.group(2).split(","):
            name = name.strip()
            arr = re.match(r"(\w+)\[3\]", name)
            c_fields.append((arr.group(1), ctype, 3) if arr else (name, ctype, 1))
    jl = (root / "julia" / "GSKrige.jl").read_text()
    jbody = re.search(r"struct GskProblem\n(.*?)\nend", jl, re.S).group(1)
    width = {"Int32": "int32_t", "Int64": "int64_t", "UInt32": "uint32_t", "Float64": "double", "Ptr{Float64}": "const double *",
             "Ptr{Int64}": "const int64_t *"}
    j_fields = []
    for line in jbody.splitlines():
        name, typ = [t.strip() for t in line.split("::")]
        tup = re.match(r"NTuple\{3,(.*)\}", typ)
        j_fields.append((name, width[tup.group(1)], 3) if tup else (name, width[typ], 1))
    assert j_fields == c_fields
    # and the ctypes mirror used by the Python host agrees on the names and order as well
    assert [f[0] for f in gsk._abi.GskProblem._fields_] == [f[0] for f in c_fields]
