"""Host-side mirror of the reference's Julia API for the DLRA hot path (Julia is not available in the build image, so
the host side above the C ABI is Python; julia/DLRAB200.jl holds the `ccall` binding a maintainer would add).

Same names, argument meaning and error behaviour as the reference:
  MatrixDEProblem / MatrixDataProblem / DLRSolution / DLRIntegrator / solve      src/primitives.jl:13-104
  ProjectorSplitting(PrimalLieTrotter|DualLieTrotter|Strang)                     src/integrators/projector_splitting.jl:1-41
  UnconventionalAlgorithm                                                        src/integrators/unconventional.jl:13-21
  RankAdaptiveUnconventionalAlgorithm(tol; rmax)                                 src/integrators/rank_adaptive_unconventional.jl:15-23
  GreedyIntegrator (SVDLike / TwoFactor data problems, MatrixHybridProblem)      src/integrators/greedy_integrator.jl:16-22,72-104
  MatrixHybridProblem, normal_component                                          src/primitives.jl:36-41, src/utils.jl:2-20
  SVDLikeRepresentation / TwoFactorRepresentation / truncated_svd                LowRankArithmetic (third party)
Everything numeric in a step happens inside libdlra.so on the GPU; this file only sequences C-ABI calls.
User right-hand sides are given in device-evaluable form (rhs.py) instead of Julia closures."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from . import _lib as L
from .engine import Engine, _is_torch

# ------------------------------------------------------------------------------------------------
# factor containers
# ------------------------------------------------------------------------------------------------


class SVDLikeRepresentation:
    """u = U*S*V' (LowRankArithmetic.SVDLikeRepresentation): host copies of the factors."""

    def __init__(self, U, S, V):
        self.U = np.array(_host(U), dtype=np.float64, order="F")
        self.S = np.array(_host(S), dtype=np.float64, order="F")
        self.V = np.array(_host(V), dtype=np.float64, order="F")

    def full(self):  # Matrix(u)
        return self.U @ self.S @ self.V.T

    @property
    def rank(self):
        return self.S.shape[0]

    @property
    def shape(self):
        return (self.U.shape[0], self.V.shape[0])

    def copy(self):
        return SVDLikeRepresentation(self.U, self.S, self.V)


class TwoFactorRepresentation:
    """u = U*Z' (LowRankArithmetic.TwoFactorRepresentation)."""

    def __init__(self, U, Z):
        self.U = np.array(_host(U), dtype=np.float64, order="F")
        self.Z = np.array(_host(Z), dtype=np.float64, order="F")

    def full(self):
        return self.U @ self.Z.T

    @property
    def rank(self):
        return self.U.shape[1]

    @property
    def shape(self):
        return (self.U.shape[0], self.Z.shape[0])

    def copy(self):
        return TwoFactorRepresentation(self.U, self.Z)


def _host(x):
    return x.detach().cpu().numpy() if _is_torch(x) else np.asarray(x)


def truncate_to_tolerance(sigma, tol) -> int:
    """LowRankArithmetic.truncate_to_tolerance; same (UNVERIFIED third-party) rule as csrc/jacobi.cuh."""
    s, r = 0.0, len(sigma)
    for sg in np.asarray(sigma)[::-1]:
        s += float(sg) ** 2
        if s > tol * tol:
            break
        r -= 1
    return r


def truncated_svd(A, r: Optional[int] = None, tol: Optional[float] = None) -> SVDLikeRepresentation:
    """LowRankArithmetic.truncated_svd(A, r) / (A; tol): initial condition helper, host LAPACK like the reference
    (outside the per-step hot path)."""
    U, s, Vt = np.linalg.svd(np.asarray(_host(A), dtype=np.float64), full_matrices=False)
    if r is None:
        r = max(1, truncate_to_tolerance(s, tol))
    return SVDLikeRepresentation(U[:, :r], np.diag(s[:r]), Vt[:r, :].T)


def truncated_svd_device(A, r: Optional[int] = None, tol: Optional[float] = None, rmax: int = 128, oversample: int = 8,
                         power_iters: int = 2, seed: int = 0, device=None) -> SVDLikeRepresentation:
    """truncated_svd for a matrix that lives on the GPU (column-major CUDA tensor): randomized subspace iteration with the
    engine's streaming kernels instead of a host LAPACK SVD of n x m.  Approximates the reference's
    exact truncation (within a few % of the optimal error for decaying spectra; exact for rank(A) <= r)."""
    n, m = A.shape
    cap = int(min(rmax if r is None else r, m, 128))
    eng = Engine(n, m, 1, cap, device=device)
    try:
        eng.truncated_svd(A, r or 0, tol or 0.0, oversample, power_iters, seed)
        return SVDLikeRepresentation(*eng.get_factors())
    finally:
        eng.close()


# ------------------------------------------------------------------------------------------------
# problems, solution, integrator  (src/primitives.jl)
# ------------------------------------------------------------------------------------------------


@dataclass
class MatrixDEProblem:
    f: object  # a device-evaluable right-hand side from rhs.py
    u0: SVDLikeRepresentation
    tspan: tuple


class MatrixDataProblem:
    def __init__(self, y, u0, tspan=None):
        self.y = y
        self.u0 = u0
        self.tspan = (1, len(y)) if tspan is None else tspan  # primitives.jl:28-30


@dataclass
class MatrixHybridProblem:  # primitives.jl:36-41
    """y(t) ~ U(t) Z(t)' with dZ/dt = f(Z, U, t).  `f` is a device-evaluable right-hand side F from rhs.py and stands for the
    Galerkin coefficient dynamics FZ(Z, U, t) = F(U Z')' U (the only form the reference uses, test/data_informed_approximation.jl:75)."""
    y: object
    f: object
    u0: "TwoFactorRepresentation"
    tspan: tuple


@dataclass
class DLRSolution:
    Y: list
    t: list


class DLRIntegrator:
    def __init__(self, engine, t, dt, sol, alg, probType, prob, save_everystep):
        self.cache = engine  # the reference's alg cache == the engine's device workspaces
        self._pushed = 0     # snapshots already handed to the engine beyond the current time (0, 1 or 2)
        self.t, self.dt, self.sol, self.alg, self.probType, self.iter = t, dt, sol, alg, probType, 0
        self.prob = prob
        self.save_everystep = save_everystep
        self.two_factor = isinstance(prob.u0, TwoFactorRepresentation)

    @property
    def u(self):
        U, S, V = self.cache.get_factors()
        if self.two_factor:   # the engine keeps Z in the V slot and S = I
            return TwoFactorRepresentation(U, V)
        return SVDLikeRepresentation(U, S, V)

    def checkpoint(self):
        """Everything needed to resume (engine state = (U,S,V,r,t,iter) + the data stream position)."""
        U, S, V = self.cache.get_factors()
        return {"U": U, "S": S, "V": V, "t": self.t, "iter": self.iter}


class _PendingSave:
    def __init__(self, fac, two_factor):
        self.fac, self.two_factor = fac, two_factor

    def resolve(self):
        U, S, V = self.fac
        return TwoFactorRepresentation(U, V) if self.two_factor else SVDLikeRepresentation(U, S, V)


def finish_saves(integ):
    """Wait for the asynchronous factor snapshots and turn the pending entries of integ.sol.Y into representations."""
    integ.cache.save_wait()
    integ.sol.Y[:] = [y.resolve() if isinstance(y, _PendingSave) else y for y in integ.sol.Y]


def init_sol(dt, t0, tf, u0):  # primitives.jl:92-104
    if isinstance(dt, (int, np.integer)) and not isinstance(dt, bool):
        steps = list(range(t0, tf + 1, dt))
        return DLRSolution([None] * len(steps), steps)
    n = int(math.floor((tf - t0) / dt)) + 1
    return DLRSolution([None] * n, list(np.linspace(t0, tf, n)))


def update_sol(integ):  # primitives.jl:82-90
    if integ.save_everystep and getattr(integ, "async_save", False):
        # the deep copy of u goes through pinned host memory on the copy stream while the next steps run; entries are
        # (U, S, V) views until solve() / finish_saves() has waited for the copies and wrapped them
        u = _PendingSave(integ.cache.save_factors_async(), integ.two_factor)
    else:
        u = integ.u if integ.save_everystep else None
    if integ.iter <= len(integ.sol.Y) - 1:
        integ.sol.Y[integ.iter] = u
        integ.sol.t[integ.iter] = integ.t
    else:
        integ.sol.Y.append(u)
        integ.sol.t.append(integ.t)


# ------------------------------------------------------------------------------------------------
# algorithms
# ------------------------------------------------------------------------------------------------


class PrimalLieTrotter:
    pass


class DualLieTrotter:
    pass


class Strang:
    pass


@dataclass
class SubStepper:
    """K_alg / S_alg / L_alg: 'tsit5' (adaptive, the reference default Tsit5()), 'tsit5_fixed', 'rk4', 'euler'."""
    kind: str = "tsit5"
    nsub: int = 1
    abstol: float = 1e-6
    reltol: float = 1e-3
    maxiters: int = 100000   # OrdinaryDiffEq default; the step raises DLRAError(EMAXITERS) beyond it


_ODE = {"euler": L.ODE_EULER, "rk4": L.ODE_RK4, "tsit5_fixed": L.ODE_TSIT5_FIXED, "tsit5": L.ODE_TSIT5}


@dataclass
class ProjectorSplitting:
    order: object = field(default_factory=PrimalLieTrotter)
    S_alg: Optional[SubStepper] = None
    L_alg: Optional[SubStepper] = None
    K_alg: Optional[SubStepper] = None


@dataclass
class UnconventionalAlgorithm:
    S_alg: Optional[SubStepper] = None
    L_alg: Optional[SubStepper] = None
    K_alg: Optional[SubStepper] = None


@dataclass
class RankAdaptiveUnconventionalAlgorithm:
    tol: float = 1e-8
    rmax: int = 2 ** 62
    S_alg: Optional[SubStepper] = None
    L_alg: Optional[SubStepper] = None
    K_alg: Optional[SubStepper] = None


@dataclass
class GreedyIntegrator:
    """greedy_integrator.jl:16-22.  Z_alg: sub-stepper of the hybrid Z-flow (default adaptive Tsit5 like the reference);
    fsal_carry: keep the Z integrator's cached first stage across steps like the reference's never-`set_u!`-ed ZIntegrator."""
    Z_alg: Optional[SubStepper] = None
    fsal_carry: bool = True


# ------------------------------------------------------------------------------------------------
# init / step! / solve
# ------------------------------------------------------------------------------------------------


def _fetch(y, t, dt):
    """update_data! (data_integrator.jl:22-28)."""
    if callable(y):
        return y(t + dt)
    if not (isinstance(t, (int, np.integer)) and isinstance(dt, (int, np.integer))):
        raise TypeError("MethodError: update_data!(x, y::AbstractArray, t::Int, dt::Int) needs integer t and dt")
    return y[t + dt - 1]


def _has_snapshot(y, tspan, t, ahead):
    """Is y(t + ahead) / y[t + ahead] inside the problem's time span?"""
    if callable(y):
        return (t + ahead) <= tspan[1] * (1 + 1e-12) + 1e-300
    return isinstance(t, (int, np.integer)) and (t + ahead - 1) < len(y) and (t + ahead) <= tspan[1]


def init(prob, alg, dt, *, device=None, comm=None, save_everystep=True, force_generic=False, lookahead=True, resume=None,
         aug_basis_first=False, async_save=False) -> DLRIntegrator:
    """init(prob, alg, dt): projector_splitting.jl:107-115, unconventional.jl:109-119,
    rank_adaptive_unconventional.jl:94-104, greedy_integrator.jl:49-59.  `comm` = "torch" (use the initialised torch.distributed group to distribute
    a fresh ncclUniqueId) or (nranks, rank, unique_id) row-shards the problem: every rank passes ITS row block of u0.U
    and of the snapshots.  A unique id can initialise ONE communicator only."""
    t0, tf = prob.tspan
    assert tf > t0, "Integration in reverse time direction is not supported"
    u0 = prob.u0
    two_factor = isinstance(u0, TwoFactorRepresentation)
    if two_factor and not isinstance(alg, GreedyIntegrator):
        raise TypeError(f"MethodError: no alg_cache for {type(alg).__name__} with a TwoFactorRepresentation")
    if isinstance(prob, MatrixHybridProblem) and not (two_factor and isinstance(alg, GreedyIntegrator)):
        raise TypeError("MethodError: MatrixHybridProblem is solved by the GreedyIntegrator on a TwoFactorRepresentation")
    if resume is not None:   # continue from DLRIntegrator.checkpoint(): factors and time come from the saved state
        u0 = (TwoFactorRepresentation(resume["U"], resume["V"]) if two_factor
              else SVDLikeRepresentation(resume["U"], resume["S"], resume["V"]))
        t0 = resume["t"]
    n, r0 = u0.U.shape
    m = u0.shape[1]
    adaptive = isinstance(alg, RankAdaptiveUnconventionalAlgorithm)
    rmax = r0
    if adaptive:
        rmax = int(min(alg.rmax, 128, m // 2 if m >= 2 else 1))
        rmax = max(rmax, r0)
        if alg.rmax < 2 ** 62 and rmax < alg.rmax:   # an explicit r_max the engine cannot honour (workspaces are sized once for 2*rmax <= 256 columns)
            import warnings
            warnings.warn(f"RankAdaptiveUnconventionalAlgorithm: rmax = {alg.rmax} clamped to {rmax} "
                          f"(libdlra.so supports augmented bases of up to 256 columns and 2*rmax <= m)")
    eng = (Engine(n, m, r0, rmax, rank_adaptive=adaptive, device=device, force_generic=force_generic, aug_basis_first=True)
           if aug_basis_first else Engine(n, m, r0, rmax, rank_adaptive=adaptive, device=device, force_generic=force_generic))
    if comm == "torch":  # torch.distributed only hands out the IPC handles / the ncclUniqueId
        from .distributed import attach_engine
        attach_engine(eng)
    elif comm is not None and comm[0] > 1:
        eng.comm_init(*comm)
    if two_factor:
        eng.set_factors(u0.U, np.eye(r0), u0.Z)  # (U, I, Z): the engine keeps Z in the V slot
    else:
        eng.set_factors(u0.U, u0.S, u0.V)  # deepcopy(prob.u0)
    if isinstance(prob, MatrixDataProblem):
        y = prob.y
        eng.data_init(y(t0) if callable(y) else y[int(t0 - prob.tspan[0])])  # yprev = y[1] | y(t0) (snapshot at the start time)
    elif isinstance(prob, MatrixHybridProblem):  # greedy_integrator.jl:41-47: ZIntegrator = init(ODEProblem(f, Z, tspan, U), Z_alg)
        prob.f.install(eng)
        sub = alg.Z_alg or SubStepper()
        eng.set_substepper(L.FLOW_L, _ODE[sub.kind], sub.nsub, sub.abstol, sub.reltol, sub.maxiters)
    else:
        if isinstance(alg, GreedyIntegrator):
            raise TypeError("MethodError: GreedyIntegrator is defined for data problems")
        prob.f.install(eng)
        for flow, sub in ((L.FLOW_K, alg.K_alg), (L.FLOW_S, alg.S_alg), (L.FLOW_L, alg.L_alg)):
            sub = sub or SubStepper()
            eng.set_substepper(flow, _ODE[sub.kind], sub.nsub, sub.abstol, sub.reltol, sub.maxiters)
    sol = init_sol(dt, t0, tf, u0)
    sol.Y[0] = u0.copy() if hasattr(u0, "copy") else u0
    integ = DLRIntegrator(eng, t0, dt, sol, alg, type(prob), prob, save_everystep)
    integ.lookahead = bool(lookahead)
    integ.async_save = bool(async_save)   # update_sol! through dlra_save_factors_async (pinned host buffers, copy stream)
    if resume is not None:
        integ.iter = 0   # sol restarts at the checkpoint; resume["iter"] tells the caller where that was
    return integ


def _push_current(integ, eng, y, t, dt):
    """update_data! for the step at time t unless a previous BUG step with lookahead already handed y(t + dt) to the engine
    (step(integ, other_alg) after a lookahead step must not push the same snapshot twice)."""
    if integ._pushed > 0:
        integ._pushed -= 1
        return
    eng.data_push(_fetch(y, t, dt))


def step(integ: DLRIntegrator, alg=None, dt=None):
    """step!(integrator, alg, dt): projector_splitting.jl:191-211, unconventional.jl:159-164,
    rank_adaptive_unconventional.jl:171-180, greedy_integrator.jl:106-111."""
    alg = integ.alg if alg is None else alg
    dt = integ.dt if dt is None else dt
    eng, t = integ.cache, integ.t
    is_data = integ.probType is MatrixDataProblem
    y = integ.prob.y if is_data else None
    if isinstance(alg, ProjectorSplitting):
        if isinstance(alg.order, Strang):
            if is_data:  # each half fetches its own increment (projector_splitting.jl:205-211)
                if integ._pushed > 0:
                    raise RuntimeError("a previous unconventional step with lookahead already handed y(t + dt) to the engine; the Strang "
                                       "step needs y(t + dt/2) next: use init(..., lookahead=False) when mixing these integrators")
                eng.data_push(_fetch(y, t, dt / 2))
                eng.step_ksl(L.KSL_PRIMAL, t, dt / 2)
                eng.data_push(_fetch(y, t + dt / 2, dt / 2))
                eng.step_ksl(L.KSL_DUAL, t + dt / 2, dt / 2)
            else:
                eng.step_ksl(L.KSL_STRANG, t, dt)
        else:
            if is_data:
                _push_current(integ, eng, y, t, dt)
            eng.step_ksl(L.KSL_PRIMAL if isinstance(alg.order, PrimalLieTrotter) else L.KSL_DUAL, t, dt)
    elif isinstance(alg, UnconventionalAlgorithm):
        if is_data:
            # update_data! for this step, plus one snapshot of lookahead when the stream has it: the engine then forms the
            # next step's K/L contractions in the same sweep as this step's core pass (identical results, 3/4 of the HBM reads)
            if integ._pushed == 0:
                eng.data_push(_fetch(y, t, dt))
                integ._pushed = 1
            if integ._pushed == 1 and integ.lookahead and _has_snapshot(y, integ.prob.tspan, t, 2 * dt):
                eng.data_push(_fetch(y, t, 2 * dt))
                integ._pushed = 2
            integ._pushed -= 1
        eng.step_bug(t, dt)
    elif isinstance(alg, RankAdaptiveUnconventionalAlgorithm):
        if is_data:
            _push_current(integ, eng, y, t, dt)
        r_new, changed = eng.step_rabug(alg.tol, alg.rmax, t, dt)
        if changed:
            print(f"rank adjusted: new rank = {r_new}")  # rank_adaptive_unconventional.jl:230
    elif isinstance(alg, GreedyIntegrator):  # greedy_step! dispatches on (typeof(u), probType), greedy_integrator.jl:72-104
        _push_current(integ, eng, integ.prob.y, t, dt)
        if integ.probType is MatrixHybridProblem:
            eng.step_greedy_two_factor(L.GREEDY_HYBRID, t, dt, alg.fsal_carry)
        elif integ.two_factor:
            eng.step_greedy_two_factor(L.GREEDY_DATA, t, dt)
        else:
            eng.step_greedy(t, dt)
    else:
        raise TypeError(f"MethodError: no step! for {type(alg).__name__}")
    integ.t += dt
    integ.iter += 1


def normal_component(integ_or_engine, dY, C=None, tol=1e-8, want_matrix=False):
    """normal_component(LRA, [C,] dY; tol) (utils.jl:2-20) for the factors currently held by the engine: the part of the
    dynamics dY that the tangent space of the low rank manifold cannot represent (a rank-adaptation indicator).  dY: device
    n x m matrix.  Returns ‖N‖_F, and N itself as a device matrix when want_matrix."""
    eng = integ_or_engine.cache if isinstance(integ_or_engine, DLRIntegrator) else integ_or_engine
    return eng.normal_component(dY, C, tol, want_matrix)


def solve(prob, alg, dt=None, **kw) -> DLRSolution:
    """LowRankIntegrators.solve (primitives.jl:68-80)."""
    if dt is None:
        assert isinstance(prob, MatrixDataProblem) and not callable(prob.y), (
            "If the data is not provided as array, integration stepsize needs to be specified")
        dt = 1
    integ = init(prob, alg, dt, **kw)
    T = prob.tspan[1] - prob.tspan[0]
    while (prob.tspan[1] - integ.t) / T > 1e-8:
        step(integ, alg, dt)
        update_sol(integ)
    if not integ.save_everystep:
        integ.sol.Y[-1] = integ.u
    integ.cache.sync()
    if integ.async_save:
        finish_saves(integ)
    return integ.sol
