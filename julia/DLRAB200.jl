# DLRAB200.jl — the `ccall` binding a LowRankIntegrators.jl maintainer would add to route the per-step DLRA hot path
# through libdlra.so (include/dlra.h).  It keeps the package's API surface: MatrixDEProblem / MatrixDataProblem,
# `LowRankIntegrators.solve(prob, alg, dt)`, ProjectorSplitting / UnconventionalAlgorithm /
# RankAdaptiveUnconventionalAlgorithm / GreedyIntegrator (also on TwoFactorRepresentation and MatrixHybridProblem) and the
# SVDLikeRepresentation factors; only `alg_cache` and `step!` change.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image (SURVEY.md F2); the same C symbols are
# exercised through ctypes by lowrankintegrators.jl_b200/_lib.py and the tests.  No CUDA.jl kernels, no CPU fallback.
#
# Seam (SURVEY.md §8b): the reference dispatches
#     alg_cache(prob, alg, u, dt; t0)                      projector_splitting.jl:43,87 | unconventional.jl:39,86 |
#                                                          rank_adaptive_unconventional.jl:47,106
#     step!(integrator::DLRIntegrator, alg, dt)            projector_splitting.jl:191-211 | unconventional.jl:159-164 |
#                                                          rank_adaptive_unconventional.jl:171-180
# A problem opts in by wrapping its algorithm:  solve(prob, OnB200(UnconventionalAlgorithm()), dt).
module DLRAB200

using LowRankIntegrators, LowRankArithmetic, LinearAlgebra
import LowRankIntegrators: alg_cache, step!, init, update_sol!, init_sol, DLRIntegrator, DLRSolution,
                           MatrixDataProblem, MatrixDEProblem, MatrixHybridProblem, AbstractDLRAlgorithm,
                           AbstractDLRAlgorithm_Cache, ProjectorSplitting, PrimalLieTrotter, DualLieTrotter, Strang,
                           UnconventionalAlgorithm, RankAdaptiveUnconventionalAlgorithm, GreedyIntegrator

const libdlra = get(ENV, "LIBDLRA", "libdlra.so")
const Handle = Ptr{Cvoid}

const DLRA_RANK_ADAPTIVE = Cint(1)
const KSL_PRIMAL, KSL_DUAL, KSL_STRANG = Cint(0), Cint(1), Cint(2)
const DATA_SNAPSHOT, DATA_DELTA = Cint(0), Cint(1)
const FLOW_K, FLOW_S, FLOW_L = Cint(0), Cint(1), Cint(2)
const GREEDY_DATA, GREEDY_HYBRID = Cint(0), Cint(1)
const ODE_EULER, ODE_RK4, ODE_TSIT5_FIXED, ODE_TSIT5 = Cint(0), Cint(1), Cint(2), Cint(3)
const OP_NONE, OP_DENSE, OP_CSR, OP_IDENTITY_SCALED = Cint(0), Cint(1), Cint(2), Cint(3)

struct DLRAError <: Exception
    code::Cint
    msg::String
end

function check(h::Handle, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:dlra_last_error, libdlra), Cstring, (Handle,), h))
    # DLRA_EINVAL mirrors the reference's MethodError / AssertionError behaviour
    throw(DLRAError(rc, msg))
end

"struct dlra_operator (include/dlra.h)"
struct Operator
    kind::Cint
    rows::Int64
    cols::Int64
    dense::Ptr{Float64}
    ld::Int64
    rowptr::Ptr{Int64}
    colind::Ptr{Int32}
    values::Ptr{Float64}
    scale::Float64
end
Operator() = Operator(OP_NONE, 0, 0, C_NULL, 0, C_NULL, C_NULL, C_NULL, 1.0)

"Device-evaluable right-hand side  F(X,t) = A·X + X·Bᵀ + G·Hᵀ + c·(D1·X).*(D2·X)  (device pointers, see dlra_rhs_set)."
Base.@kwdef struct FactoredRHS
    A::Operator = Operator()
    B::Operator = Operator()
    G::Ptr{Float64} = C_NULL
    ldg::Int64 = 0
    H::Ptr{Float64} = C_NULL
    ldh::Int64 = 0
    q::Cint = 0
    D1::Operator = Operator()
    D2::Operator = Operator()
    c_had::Float64 = 0.0
end

"Algorithm wrapper that selects the B200 engine; `device` is the CUDA ordinal of this process."
struct OnB200{A<:AbstractDLRAlgorithm} <: AbstractDLRAlgorithm
    alg::A
    device::Cint
end
OnB200(alg) = OnB200(alg, Cint(0))

"The engine handle plays the role of the reference's alg cache (all workspaces live on the device)."
mutable struct B200Cache <: AbstractDLRAlgorithm_Cache
    h::Handle
    y            # data stream of a MatrixDataProblem (snapshots on the host or device pointers), else nothing
    n::Int
    m::Int
    pushed::Int  # snapshots already handed to the engine beyond the current time (0, 1 or 2: one snapshot of lookahead)
    tf           # end of the time span (no lookahead past it)
    two_factor::Bool  # u = U*Z': the engine keeps Z in the V slot and S = I
end

function B200Cache(device, n, m, r0, rmax, adaptive::Bool)
    href = Ref{Handle}(C_NULL)
    rc = ccall((:dlra_create, libdlra), Cint, (Cint, Int64, Int64, Cint, Cint, Cint, Ref{Handle}),
               device, n, m, r0, rmax, adaptive ? DLRA_RANK_ADAPTIVE : Cint(0), href)
    rc == 0 || throw(DLRAError(rc, unsafe_string(ccall((:dlra_last_error, libdlra), Cstring, (Handle,), C_NULL))))
    c = B200Cache(href[], nothing, n, m, 0, nothing, false)
    finalizer(x -> ccall((:dlra_destroy, libdlra), Cint, (Handle,), x.h), c)
    return c
end

set_factors!(c::B200Cache, u::SVDLikeRepresentation) =
    check(c.h, ccall((:dlra_set_factors_host, libdlra), Cint,
                     (Handle, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cint),
                     c.h, u.U, size(u.U, 1), u.S, size(u.S, 1), u.V, size(u.V, 1), rank(u)))

function set_factors!(c::B200Cache, u::TwoFactorRepresentation)
    r = size(u.U, 2)
    c.two_factor = true
    check(c.h, ccall((:dlra_set_factors_host, libdlra), Cint,
                     (Handle, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cint),
                     c.h, u.U, size(u.U, 1), Matrix{Float64}(I, r, r), r, u.Z, size(u.Z, 1), r))
end

"update_sol! (primitives.jl:82-90): deep copy of the device factors into a fresh SVDLikeRepresentation"
function get_factors(c::B200Cache)
    r = Ref{Cint}(0)
    check(c.h, ccall((:dlra_get_rank, libdlra), Cint, (Handle, Ref{Cint}), c.h, r))
    U, S, V = zeros(c.n, r[]), zeros(r[], r[]), zeros(c.m, r[])
    check(c.h, ccall((:dlra_get_factors_host, libdlra), Cint,
                     (Handle, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ref{Cint}),
                     c.h, U, c.n, S, r[], V, c.m, r))
    return c.two_factor ? TwoFactorRepresentation(U, V) : SVDLikeRepresentation(U, S, V)
end

ode_code(alg) = alg isa Tsit5 ? ODE_TSIT5 : alg isa RK4 ? ODE_RK4 : alg isa Euler ? ODE_EULER :
                throw(ArgumentError("sub-stepper $(typeof(alg)) has no device counterpart (Tsit5, RK4, Euler)"))

# ---- alg_cache ---------------------------------------------------------------------------------------------------
function alg_cache(prob::MatrixDataProblem, w::OnB200, u, dt; t0 = prob.tspan[1])
    n, r = size(u.U); m = size(u, 2)
    adaptive = w.alg isa RankAdaptiveUnconventionalAlgorithm
    rmax = adaptive ? Int(min(w.alg.alg_params.r_max, 128, m ÷ 2)) : r
    c = B200Cache(w.device, n, m, r, max(rmax, r), adaptive)
    set_factors!(c, u)
    c.y = prob.y
    c.tf = prob.tspan[2]
    y0 = prob.y isa AbstractArray ? prob.y[1] : prob.y(t0)          # yprev (projector_splitting.jl:91)
    check(c.h, ccall((:dlra_data_init_host, libdlra), Cint, (Handle, Ptr{Float64}, Int64), c.h, y0, size(y0, 1)))
    return c
end

function alg_cache(prob::MatrixDEProblem{<:FactoredRHS}, w::OnB200, u, dt; t0 = prob.tspan[1])
    n, r = size(u.U); m = size(u.V, 1)
    adaptive = w.alg isa RankAdaptiveUnconventionalAlgorithm
    rmax = adaptive ? Int(min(w.alg.alg_params.r_max, 128, m ÷ 2)) : r
    c = B200Cache(w.device, n, m, r, max(rmax, r), adaptive)
    set_factors!(c, u)
    f = prob.f
    check(c.h, ccall((:dlra_rhs_set, libdlra), Cint,
                     (Handle, Ref{Operator}, Ref{Operator}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cint,
                      Ref{Operator}, Ref{Operator}, Float64),
                     c.h, f.A, f.B, f.G, f.ldg, f.H, f.ldh, f.q, f.D1, f.D2, f.c_had))
    p = w.alg.alg_params
    for (flow, a, kw) in ((FLOW_K, p.K_alg, p.K_kwargs), (FLOW_S, p.S_alg, p.S_kwargs), (FLOW_L, p.L_alg, p.L_kwargs))
        check(c.h, ccall((:dlra_set_substepper, libdlra), Cint, (Handle, Cint, Cint, Cint, Float64, Float64),
                         c.h, flow, ode_code(a), Cint(get(kw, :nsub, 1)), get(kw, :abstol, 1e-6), get(kw, :reltol, 1e-3)))
    end
    return c
end

# greedy_integrator.jl:41-47: the Z-flow dZ/dt = F(U Z')' U (FZ of test/data_informed_approximation.jl:75) runs on the device
function alg_cache(prob::MatrixHybridProblem{<:Any,<:FactoredRHS}, w::OnB200{<:GreedyIntegrator}, u::TwoFactorRepresentation, dt;
                   t0 = prob.tspan[1])
    n, r = size(u.U); m = size(u.Z, 1)
    c = B200Cache(w.device, n, m, r, r, false)
    set_factors!(c, u)
    c.y = prob.y
    c.tf = prob.tspan[2]
    f = prob.f
    check(c.h, ccall((:dlra_rhs_set, libdlra), Cint,
                     (Handle, Ref{Operator}, Ref{Operator}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cint,
                      Ref{Operator}, Ref{Operator}, Float64),
                     c.h, f.A, f.B, f.G, f.ldg, f.H, f.ldh, f.q, f.D1, f.D2, f.c_had))
    p = w.alg.alg_params
    check(c.h, ccall((:dlra_set_substepper, libdlra), Cint, (Handle, Cint, Cint, Cint, Float64, Float64),
                     c.h, FLOW_L, ode_code(p.Z_alg), Cint(get(p.Z_kwargs, :nsub, 1)), get(p.Z_kwargs, :abstol, 1e-6),
                     get(p.Z_kwargs, :reltol, 1e-3)))
    return c
end

"""
    normal_component(cache, dY_device, ld; tol = 1e-8) -> ‖(I-UU')·dY·(I-Z·pinv(Z'Z, atol=tol)·Z')‖_F

utils.jl:2-20 for the factors held by the engine (dY: device pointer to the n_local x m dynamics, e.g. a CuArray).
"""
function normal_component(c::B200Cache, dY::Ptr{Float64}, ld::Integer; tol = 1e-8)
    nrm = Ref{Float64}(0.0)
    check(c.h, ccall((:dlra_normal_component, libdlra), Cint,
                     (Handle, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Ptr{Float64}, Int64, Ref{Float64}),
                     c.h, dY, ld, C_NULL, 0, tol, C_NULL, 0, nrm))
    return nrm[]
end

"F(X) += A*X*B' (two-sided term of the installed right-hand side; see dlra_rhs_add_term)"
add_term!(c::B200Cache, A::Operator, B::Operator) =
    check(c.h, ccall((:dlra_rhs_add_term, libdlra), Cint, (Handle, Ref{Operator}, Ref{Operator}), c.h, A, B))

function init(prob, w::OnB200, dt)
    t0, tf = prob.tspan
    @assert tf > t0 "Integration in reverse time direction is not supported"
    u = deepcopy(prob.u0)
    sol = init_sol(dt, t0, tf, prob.u0)
    cache = alg_cache(prob, w, u, dt, t0 = t0)
    sol.Y[1] = deepcopy(prob.u0)
    return DLRIntegrator(u, t0, dt, sol, w, cache, typeof(prob), 0)
end

# ---- step! -------------------------------------------------------------------------------------------------------
push_data!(c::B200Cache, t, dt) = begin
    y = c.y isa AbstractArray ? c.y[t + dt] : c.y(t + dt)           # update_data! (data_integrator.jl:22-28)
    check(c.h, ccall((:dlra_data_push_host, libdlra), Cint, (Handle, Ptr{Float64}, Int64, Cint), c.h, y, size(y, 1), DATA_SNAPSHOT))
end

ksl!(c, order, t, dt) = check(c.h, ccall((:dlra_step_ksl, libdlra), Cint, (Handle, Cint, Float64, Float64), c.h, order, t, dt))

function step!(integrator::DLRIntegrator, w::OnB200, dt)
    c, t, alg = integrator.cache, integrator.t, w.alg
    isdata = integrator.probType <: MatrixDataProblem
    if alg isa ProjectorSplitting
        if alg.order isa Strang
            if isdata
                push_data!(c, t, dt / 2);          ksl!(c, KSL_PRIMAL, t, dt / 2)
                push_data!(c, t + dt / 2, dt / 2); ksl!(c, KSL_DUAL, t + dt / 2, dt / 2)
            else
                ksl!(c, KSL_STRANG, t, dt)
            end
        else
            isdata && push_data!(c, t, dt)
            ksl!(c, alg.order isa PrimalLieTrotter ? KSL_PRIMAL : KSL_DUAL, t, dt)
        end
    elseif alg isa UnconventionalAlgorithm
        if isdata
            # update_data! for this step plus one snapshot of lookahead: libdlra then forms the next step's K/L contractions
            # in the same sweep as this step's core pass (identical results, 3/4 of the HBM reads)
            c.pushed == 0 && (push_data!(c, t, dt); c.pushed = 1)
            if c.pushed == 1 && t + 2dt <= c.tf
                push_data!(c, t, 2dt); c.pushed = 2
            end
            c.pushed -= 1
        end
        check(c.h, ccall((:dlra_step_bug, libdlra), Cint, (Handle, Float64, Float64), c.h, t, dt))
    elseif alg isa RankAdaptiveUnconventionalAlgorithm
        isdata && push_data!(c, t, dt)
        rnew, changed = Ref{Cint}(0), Ref{Cint}(0)
        check(c.h, ccall((:dlra_step_rabug, libdlra), Cint, (Handle, Float64, Float64, Float64, Int64, Ref{Cint}, Ref{Cint}),
                         c.h, t, dt, alg.alg_params.tol, min(alg.alg_params.r_max, typemax(Int64)), rnew, changed))
        changed[] != 0 && println("rank adjusted: new rank = $(rnew[])")   # rank_adaptive_unconventional.jl:230
    elseif alg isa GreedyIntegrator                                         # greedy_step! methods, greedy_integrator.jl:72-104
        push_data!(c, t, dt)                                                # X = y(t+dt) | y[t+dt]
        if integrator.probType <: MatrixHybridProblem
            check(c.h, ccall((:dlra_step_greedy_two_factor, libdlra), Cint, (Handle, Cint, Cint, Float64, Float64),
                             c.h, GREEDY_HYBRID, Cint(1), t, dt))
        elseif c.two_factor
            check(c.h, ccall((:dlra_step_greedy_two_factor, libdlra), Cint, (Handle, Cint, Cint, Float64, Float64),
                             c.h, GREEDY_DATA, Cint(1), t, dt))
        else
            check(c.h, ccall((:dlra_step_greedy, libdlra), Cint, (Handle, Float64, Float64), c.h, t, dt))
        end
    else
        throw(MethodError(step!, (integrator, w, dt)))
    end
    integrator.u = get_factors(c)      # the host copy the reference mutates in place
    integrator.t += dt
    integrator.iter += 1
end

# ---- multi-GPU (one Julia process per GPU, e.g. MPI.jl or Distributed.jl workers) -----------------------------------
"""
    attach!(cache, nranks, rank, allgather)

Row-shard the problem: every process passes ITS contiguous row block of `u0.U` and of each snapshot.  `allgather(bytes)`
is any host-side all-gather of a 64-byte blob (MPI.Allgather, Distributed.jl ...): it hands out the CUDA-IPC handles of the
peer-to-peer exchange regions (dlra_p2p_export / dlra_p2p_import).  NCCL (dlra_nccl_unique_id / dlra_comm_init) is the
alternative transport.
"""
function attach!(c::B200Cache, nranks::Integer, rank::Integer, allgather)
    mine = Vector{UInt8}(undef, 64)
    check(c.h, ccall((:dlra_p2p_export, libdlra), Cint, (Handle, Ptr{UInt8}), c.h, mine))
    all = allgather(mine)::Vector{UInt8}                      # nranks * 64 bytes, rank order
    check(c.h, ccall((:dlra_p2p_import, libdlra), Cint, (Handle, Cint, Cint, Ptr{UInt8}), c.h, nranks, rank, all))
end

# ---- asynchronous update_sol! (primitives.jl:82-90) and step progress -----------------------------------------------------------
# `bufs = (U, S, V)` are host matrices the caller keeps alive (pinned with CUDA.Mem.pin for a truly asynchronous copy); they hold
# the factors of the step after which this was called once `save_wait` has returned.  The step stream is not stalled.
function save_factors_async!(c::B200Cache, U::Matrix{Float64}, S::Matrix{Float64}, V::Matrix{Float64})
    r = Ref{Cint}(0)
    check(c.h, ccall((:dlra_save_factors_async, libdlra), Cint,
                     (Handle, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ref{Cint}),
                     c.h, U, size(U, 1), S, size(S, 1), V, size(V, 1), r))
    return Int(r[])
end
save_wait(c::B200Cache) = check(c.h, ccall((:dlra_save_wait, libdlra), Cint, (Handle,), c.h))

# (steps enqueued, steps completed on the device); wait_for > completed blocks until that many steps are done
function progress(c::B200Cache; wait_for::Integer = -1)
    enq = Ref{Int64}(0); done = Ref{Int64}(0)
    check(c.h, ccall((:dlra_progress, libdlra), Cint, (Handle, Ref{Int64}, Ref{Int64}, Int64), c.h, enq, done, wait_for))
    return Int(enq[]), Int(done[])
end

# OrdinaryDiffEq's `maxiters` of the K/S/L sub-integrators (flow = FLOW_K | FLOW_S | FLOW_L); a step that hits it throws
# DLRAError(7, ...) (DLRA_EMAXITERS, the reference's retcode MaxIters) and leaves integrator.u untouched
set_maxiters!(c::B200Cache, flow::Integer, maxiters::Integer) =
    check(c.h, ccall((:dlra_set_substepper_maxiters, libdlra), Cint, (Handle, Cint, Int64), c.h, flow, maxiters))

end # module
