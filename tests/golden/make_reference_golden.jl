# Pins the oracle against the REAL reference.  Neither this repository's build image nor its GPU boxes have Julia, so the
# oracle (oracle/dlra_oracle.py) is pinned only structurally; a maintainer with Julia runs this script ONCE:
#
#     julia --project=/path/to/LowRankIntegrators.jl tests/golden/make_reference_golden.jl tests/golden/reference
#
# It writes seeded inputs and the outputs of `LowRankIntegrators.solve` (five integrators on a data problem, five on a
# MatrixDEProblem with the default adaptive Tsit5 sub-integrators) plus probes of `LowRankArithmetic.truncate_to_tolerance`
# and `truncated_svd` as raw little-endian Float64 column-major files with a manifest.  tests/test_reference_golden.py
# consumes the directory when it exists (oracle on CPU, engine on GPU) and is skipped otherwise.
#
# Reference lines exercised: src/primitives.jl:68-104, src/integrators/data_integrator.jl:1-28,
# projector_splitting.jl:87-211, unconventional.jl:86-164, rank_adaptive_unconventional.jl:106-233,
# greedy_integrator.jl:94-104; third-party: LowRankArithmetic.truncate_to_tolerance / truncated_svd, OrdinaryDiffEq Tsit5().
using LowRankIntegrators, LowRankArithmetic, LinearAlgebra, Random, Printf

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "reference")
mkpath(outdir)
manifest = IOBuffer()

function put(name::String, A::AbstractArray{Float64})
    open(joinpath(outdir, name * ".bin"), "w") do io
        write(io, htol.(reinterpret(UInt64, vec(collect(A)))))
    end
    println(manifest, name, " ", join(size(A), " "))
end
put(name::String, x::Real) = put(name, [Float64(x)])
put_rep(name, u) = (put(name * ".U", Matrix(u.U)); put(name * ".S", Matrix(u.S)); put(name * ".V", Matrix(u.V)))

Random.seed!(20261017)

# ---- 1. truncate_to_tolerance / truncated_svd probes ------------------------------------------------------------------
sigmas = [[2.0^(-j) for j in 1:20], [1.0, 1e-3, 1e-6, 1e-9, 1e-12], [3.0, 3.0, 3.0, 1e-8, 1e-8], [1.0], [0.5, 0.25, 0.0, 0.0]]
tols = [1e-1, 1e-4, 1e-8, 1e-12, 0.0, 10.0]
for (i, s) in enumerate(sigmas)
    put("ttt.sigma$(i)", s)
    put("ttt.rank$(i)", Float64[LowRankArithmetic.truncate_to_tolerance(s, tol) for tol in tols])
end
put("ttt.tols", tols)
A = randn(40, 12) * Diagonal([10.0^(-j) for j in 0:11]) * randn(12, 30)
put("tsvd.A", A)
put_rep("tsvd.r5", truncated_svd(A, 5))
put_rep("tsvd.tol1e-6", truncated_svd(A, tol = 1e-6))

# ---- 2. MatrixDataProblem: discrete snapshot stream, the five integrators --------------------------------------------
n, m, R, r, nsnap = 120, 80, 12, 6, 6
P, W = 2 .* rand(n, R) .- 1, 2 .* rand(m, R) .- 1
sig = [2.0^(-q) for q in 0:R-1]; om = 0.5 .+ 1.5 .* rand(R); ph = 2pi .* rand(R); H = 2 .* rand(n, m) .- 1
Yt(t) = (P .* (sig .* cos.(om .* t .+ ph))') * W' + 1e-4 * cos(3t) * H
snaps = [Yt(0.05 * k) for k in 0:nsnap-1]
for (k, s) in enumerate(snaps); put("data.snap$(k)", s); end
X0 = truncated_svd(snaps[1], r)
put_rep("data.u0", X0)
data_solvers = ["bug" => UnconventionalAlgorithm(), "ksl_primal" => ProjectorSplitting(PrimalLieTrotter()),
                "ksl_dual" => ProjectorSplitting(DualLieTrotter()), "rabug" => RankAdaptiveUnconventionalAlgorithm(1e-6, rmax = 12),
                "greedy" => GreedyIntegrator()]
for (name, alg) in data_solvers
    sol = LowRankIntegrators.solve(MatrixDataProblem(snaps, deepcopy(X0)), alg)
    put("data.$(name).nsol", length(sol.Y))
    for k in 2:length(sol.Y)
        put_rep("data.$(name).step$(k-1)", sol.Y[k])
    end
end
# continuous stream + Strang (two increments per step)
sol = LowRankIntegrators.solve(MatrixDataProblem(Yt, deepcopy(X0), (0.0, 0.25)), ProjectorSplitting(Strang()), 0.05)
put("data.P", P); put("data.W", W); put("data.sig", sig); put("data.om", om); put("data.ph", ph); put("data.H", H)
put("data.strang.nsol", length(sol.Y))
for k in 2:length(sol.Y); put_rep("data.strang.step$(k-1)", sol.Y[k]); end

# ---- 3. MatrixDEProblem (examples/generic_matrix.jl form, default Tsit5 sub-integrators) ------------------------------
N = 40
function skew(N)
    Wm = randn(N, N)
    for i in 1:N
        Wm[i, i] = 0
        for j in 1:i; Wm[i, j] = -Wm[j, i]; end
    end
    return Wm
end
W1, W2 = 0.2 * skew(N), 0.2 * skew(N)
put("de.W1", W1); put("de.W2", W2)
F(X, t) = W1 * X + X + X * W2
D0 = Matrix(Diagonal([2.0^(-j) for j in 1:N])) + 1e-3 * randn(N, N)
put("de.X0full", D0)
XD = truncated_svd(D0, 5)
put_rep("de.u0", XD)
de_solvers = ["bug" => UnconventionalAlgorithm(), "ksl_primal" => ProjectorSplitting(PrimalLieTrotter()),
              "ksl_dual" => ProjectorSplitting(DualLieTrotter()), "ksl_strang" => ProjectorSplitting(Strang()),
              "rabug" => RankAdaptiveUnconventionalAlgorithm(1e-5, rmax = 10)]
for (name, alg) in de_solvers
    sol = LowRankIntegrators.solve(MatrixDEProblem(F, deepcopy(XD), (0.0, 0.05)), alg, 0.01)
    put("de.$(name).nsol", length(sol.Y))
    for k in 2:length(sol.Y); put_rep("de.$(name).step$(k-1)", sol.Y[k]); end
end

open(joinpath(outdir, "manifest.txt"), "w") do io
    write(io, String(take!(manifest)))
end
println("wrote ", outdir)
