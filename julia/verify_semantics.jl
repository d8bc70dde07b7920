# verify_semantics.jl — run on a machine WITH Julia and the pinned packages (GeoStatsSolvers 0.7.16 ⇒
# GeoStatsModels 0.2, Variography 0.22, Meshes 0.37) to close the [3P-RECALLED] checklist of SURVEY.md §8c.
# Each line prints what oracle/gsk_oracle.c and libgskrige.so assume; a mismatch means flipping the
# corresponding switch in gsk_problem (support offsets, gaussian_nugget_eps, flags) — no code change.
using GeoStatsSolvers, GeoStatsModels, Variography, Meshes, GeoTables, LinearAlgebra

println("V1 block support of a unit 2-D cell (expect 9 points at {0.25,0.5,0.75}²):")
q = Quadrangle((0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0))
println(collect(Variography._sample(GaussianVariogram(range=35.0), q)))
println("V2 Gaussian nugget epsilon (expect γ(1e-9) ≈ 1e-6): ", GaussianVariogram(range=35.0)(1e-9))
println("V4 UK exponents degree 1, dim 3 (expect x,y,z,1): ", GeoStatsModels.UniversalKriging(GaussianVariogram(), 1, 3).exponents)
println("V4 UK exponents degree 2, dim 2 (expect x²,y²,x,y,xy,1): ", GeoStatsModels.UniversalKriging(GaussianVariogram(), 2, 2).exponents)
data = georef((; z=[1.0, 0.0, 1.0]), [(25.0, 25.0), (50.0, 75.0), (75.0, 50.0)])
grid = CartesianGrid((100, 100), (0.5, 0.5), (1.0, 1.0))
println("V9 centroid(grid, 1) (expect (1.0, 1.0)): ", centroid(grid, 1))
sol = solve(EstimationProblem(data, grid, :z), KrigingSolver(:z => (variogram=GaussianVariogram(range=35.0, nugget=0.0), maxneighbors=3)))
# compare with:  python -c "import gskrige, oracle_py; ..."  (tests/_cases.py: ref_problem_2d)
println("known answer: mean[1:3] = ", sol.z[1:3], "  var[1:3] = ", sol.z_variance[1:3])
println("expected from the oracle: mean[0:3] and var[0:3] of tests/_cases.ref_problem_2d(k=3)")
