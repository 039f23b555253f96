"""CPU oracle for the per-step DLRA hot path of FHoltorf/LowRankIntegrators.jl.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module;
the product path (``lowrankintegrators.jl_b200`` -> ``libdlra.so``) never does.

PARITY UNPINNED: the reference is pure Julia, Julia is not installed in the build
container or on the GPU box, and the reference's own tests hold no golden vectors
and no seeded RNG.  This file is therefore a line-by-line NumPy
(OpenBLAS/LAPACK, fp64) restatement of the reference sources, pinned only by
  (i)   the reference's own self-consistency test (continuous == discrete stream,
        test/data_driven_approximation.jl:30),
  (ii)  the exactness property of KSL/BUG on rank-r data (README.md refs [1],[2]),
  (iii) the analytic best-rank-r error of examples/generic_matrix.jl:16-18,34,
all exercised in tests/test_oracle.py.  Third-party arithmetic that is absent from
/root/reference (LowRankArithmetic >=0.1.2,<0.2 and OrdinaryDiffEq's Tsit5 via
DifferentialEquations 7; Project.toml:7-15, no Manifest) is restated from its published
algorithm and flagged UNVERIFIED where it matters (truncate_to_tolerance, the adaptive
step-size controller).

All `file:line` citations are relative to /root/reference/.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Union

import numpy as np

# ----------------------------------------------------------------------------------------------
# factor containers (LowRankArithmetic.jl, used at projector_splitting.jl:54,64,74 and
# rank_adaptive_unconventional.jl:231; README.md:85,92)
# ----------------------------------------------------------------------------------------------


class SVDLikeRepresentation:
    """u = U * S * V'  (LowRankArithmetic.SVDLikeRepresentation; mutable fields U,S,V)."""

    def __init__(self, U, S, V):
        self.U = np.array(U, dtype=np.float64, order="F")
        self.S = np.array(S, dtype=np.float64, order="F")
        self.V = np.array(V, dtype=np.float64, order="F")

    def full(self):  # Matrix(u)
        return self.U @ self.S @ self.V.T

    @property
    def rank(self):
        return self.S.shape[0]

    @property
    def shape(self):
        return (self.U.shape[0], self.V.shape[0])

    def copy(self):
        return SVDLikeRepresentation(self.U.copy(), self.S.copy(), self.V.copy())


class TwoFactorRepresentation:
    """u = U * Z'  (LowRankArithmetic.TwoFactorRepresentation)."""

    def __init__(self, U, Z):
        self.U = np.array(U, dtype=np.float64, order="F")
        self.Z = np.array(Z, dtype=np.float64, order="F")

    def full(self):
        return self.U @ self.Z.T

    @property
    def rank(self):
        return self.U.shape[1]

    @property
    def shape(self):
        return (self.U.shape[0], self.Z.shape[0])

    def copy(self):
        return TwoFactorRepresentation(self.U.copy(), self.Z.copy())


def truncate_to_tolerance(sigma, tol) -> int:
    """LowRankArithmetic.truncate_to_tolerance (call site rank_adaptive_unconventional.jl:223).

    UNVERIFIED (source not in container): smallest r such that the discarded tail
    satisfies sqrt(sum_{j>r} sigma_j^2) <= tol; accumulated from the tail, stops when the
    running sum of squares exceeds tol^2.  Kept as ONE swappable function; the CUDA engine's
    rank-selection kernel restates exactly this loop.
    """
    s = 0.0
    r = len(sigma)
    for sg in sigma[::-1]:
        s += float(sg) * float(sg)
        if s > tol * tol:
            break
        r -= 1
    return r


def truncated_svd(A, r: Optional[int] = None, tol: Optional[float] = None) -> SVDLikeRepresentation:
    """LowRankArithmetic.truncated_svd(A, r) / truncated_svd(A; tol) (call sites
    test/data_driven_approximation.jl:18, test/data_agnostic_approximation.jl:45,
    examples/generic_matrix.jl:29): LAPACK SVD truncated to r, S = Matrix(Diagonal(sigma[1:r]))."""
    U, s, Vt = np.linalg.svd(np.asarray(A, dtype=np.float64), full_matrices=False)
    if r is None:
        r = max(1, truncate_to_tolerance(s, tol))
    return SVDLikeRepresentation(U[:, :r], np.diag(s[:r]), Vt[:r, :].T)


# ----------------------------------------------------------------------------------------------
# problems / solution / driver  (src/primitives.jl)
# ----------------------------------------------------------------------------------------------


@dataclass
class MatrixDEProblem:  # primitives.jl:13-17
    f: Callable  # f(Y_dense, t) -> dense n x m   (the oracle evaluates F densely, small sizes only)
    u0: SVDLikeRepresentation
    tspan: tuple


class MatrixDataProblem:  # primitives.jl:23-30
    def __init__(self, y, u0, tspan=None):
        self.y = y
        self.u0 = u0
        if tspan is None:
            # MatrixDataProblem(y::AbstractArray, u0) = MatrixDataProblem(y, u0, (1, length(y)))
            tspan = (1, len(y))
        self.tspan = tspan


@dataclass
class MatrixHybridProblem:  # primitives.jl:36-41
    y: object      # snapshot function t -> n x m (or a sequence of snapshots)
    f: Callable    # f(Z, U, t) -> m x r : dZ/dt for the coefficients of y ~ U*Z'
    u0: object     # TwoFactorRepresentation
    tspan: tuple


@dataclass
class DLRSolution:  # primitives.jl:46-49
    Y: list
    t: list


@dataclass
class DLRIntegrator:  # primitives.jl:54-63
    u: object
    t: float
    dt: float
    sol: DLRSolution
    alg: object
    cache: object
    probType: type
    iter: int = 0


def init_sol(dt, t0, tf, u0) -> DLRSolution:
    """primitives.jl:92-104."""
    if isinstance(dt, (int, np.integer)) and not isinstance(dt, bool):
        steps = list(range(t0, tf + 1, dt))  # t0:dt:tf
        return DLRSolution([None] * len(steps), list(steps))
    n = int(math.floor((tf - t0) / dt)) + 1
    return DLRSolution([None] * n, list(np.linspace(t0, tf, n)))


def update_sol(integ: DLRIntegrator):
    """primitives.jl:82-90 (deep copy of u per step)."""
    if integ.iter <= len(integ.sol.Y) - 1:
        integ.sol.Y[integ.iter] = integ.u.copy()
        integ.sol.t[integ.iter] = integ.t
    else:
        integ.sol.Y.append(integ.u.copy())
        integ.sol.t.append(integ.t)


def solve(prob, alg, dt=None) -> DLRSolution:
    """primitives.jl:68-80."""
    if dt is None:
        assert isinstance(prob, MatrixDataProblem) and not callable(prob.y), (
            "If the data is not provided as array, integration stepsize needs to be specified")
        dt = 1
    integ = init(prob, alg, dt)
    T = prob.tspan[1] - prob.tspan[0]
    while (prob.tspan[1] - integ.t) / T > 1e-8:
        step(integ, alg, dt)
        update_sol(integ)
    return integ.sol


# ----------------------------------------------------------------------------------------------
# data sub-integrator and data feed  (src/integrators/data_integrator.jl)
# ----------------------------------------------------------------------------------------------


def update_data(y, t, dt):
    """data_integrator.jl:22-28: x .= y(t+dt)  |  x .= deepcopy(y[t+dt]) (1-based snapshot index)."""
    if callable(y):
        return np.array(y(t + dt), dtype=np.float64)
    assert isinstance(t, (int, np.integer)) and isinstance(dt, (int, np.integer)), (
        "MethodError: update_data!(x, y::AbstractArray, t::Int, dt::Int)")
    return np.array(y[t + dt - 1], dtype=np.float64)


def data_contract(dy, left, right, sign=1):
    """data_integrator.jl:13-16: sign * left' * dy * right (left/right == None stands for I)."""
    out = dy
    if left is not None:
        out = left.T @ out
    if right is not None:
        out = out @ right
    return sign * out


class _DataFeed:
    """alg_cache(::MatrixDataProblem, ...) data part (projector_splitting.jl:87-93 and twins) and the
    ΔA formation (projector_splitting.jl:117-121,154-158; unconventional.jl:121-125;
    rank_adaptive_unconventional.jl:182-186)."""

    def __init__(self, y, t0):
        self.y = y
        self.yprev = np.array(y(t0) if callable(y) else y[0], dtype=np.float64)

    def advance(self, t, dt):
        ycurr = update_data(self.y, t, dt)
        dy = ycurr - self.yprev
        self.yprev = ycurr
        return dy


# ----------------------------------------------------------------------------------------------
# explicit RK sub-steppers standing in for OrdinaryDiffEq (third party)
# ----------------------------------------------------------------------------------------------

TSIT5_C = (0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0)
TSIT5_A = (
    (),
    (0.161,),
    (-0.008480655492356989, 0.335480655492357),
    (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
    (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
    (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
    (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774),
)
TSIT5_BTILDE = (-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                0.5823571654525552, -0.45808210592918697, 0.015151515151515152)


class MaxItersError(RuntimeError):
    """OrdinaryDiffEq returns retcode MaxIters (with a warning) when `maxiters` sub-steps do not reach the requested time."""


@dataclass
class SubStepper:
    """Stand-in for the `*_alg` / `*_kwargs` of the algorithm constructors
    (projector_splitting.jl:36-38, unconventional.jl:16-18, rank_adaptive_unconventional.jl:18-20).
    kind: 'euler' | 'rk4' | 'tsit5_fixed' | 'tsit5' (adaptive, default like the reference's Tsit5()).
    nsub: number of equal sub-steps per outer step for the fixed-step kinds."""
    kind: str = "tsit5"
    nsub: int = 1
    abstol: float = 1e-6
    reltol: float = 1e-3
    maxiters: int = 100000   # OrdinaryDiffEq default `maxiters` (the reference passes none): MaxItersError beyond it
    # adaptive controller state (carried across outer steps like an OrdinaryDiffEq integrator object)
    dt_next: Optional[float] = None
    qold: float = 1e-4
    nfev: int = 0
    naccept: int = 0
    nreject: int = 0


def _tsit5_stages(f, u, t, h, k1):
    ks = [k1]
    for s in range(1, 7):
        us = u.copy()
        for j, a in enumerate(TSIT5_A[s]):
            if a != 0.0:
                us = us + (h * a) * ks[j]
        if s < 6:
            ks.append(f(us, t + TSIT5_C[s] * h))
        else:
            unew = us
            ks.append(f(unew, t + h))
    return unew, ks


def ode_advance(stepper: SubStepper, f, u0, t0, dt, carry: Optional[dict] = None):
    """`set_u!(I, u0); step!(I, dt, true); I.u` (e.g. unconventional.jl:137-139): integrate
    u' = f(u, t) from t0 to exactly t0+dt.  Spec shared verbatim with the CUDA engine
    (csrc/de_flows.cuh); adaptive controller per SURVEY.md Appendix B (UNVERIFIED vs OrdinaryDiffEq).

    carry: None for the K/S/L integrators (their `set_u!` invalidates the cached first stage).  The greedy
    hybrid step never calls `set_u!` on its ZIntegrator (greedy_integrator.jl:72-76), so OrdinaryDiffEq
    re-uses the derivative it evaluated at the end of the previous outer step as the first stage -- evaluated
    with the basis U of *that* step, although `mul!(U, Q, P')` (:81) has changed U in place since.  `carry`
    (a dict owned by the cache) holds that derivative across calls to reproduce this.  UNVERIFIED against
    OrdinaryDiffEq like the rest of this function."""
    u = np.array(u0, dtype=np.float64)
    kind = stepper.kind

    def first_stage(t):
        if carry is not None and carry.get("k") is not None and carry["k"].shape == u.shape:
            return carry["k"]
        stepper.nfev += 1
        return f(u, t)

    if kind in ("euler", "rk4", "tsit5_fixed"):
        h = dt / stepper.nsub
        t = t0
        klast = None
        for it in range(stepper.nsub):
            if it == 0:
                k1 = first_stage(t)
            else:
                k1 = f(u, t)
                stepper.nfev += 1
            if kind == "euler":
                u = u + h * k1
            elif kind == "rk4":
                k2 = f(u + (0.5 * h) * k1, t + 0.5 * h)
                k3 = f(u + (0.5 * h) * k2, t + 0.5 * h)
                k4 = f(u + h * k3, t + h)
                u = u + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
                stepper.nfev += 3
            else:
                u, ks = _tsit5_stages(f, u, t, h, k1)
                klast = ks[6]
                stepper.nfev += 6
            t += h
        if carry is not None:
            if kind == "tsit5_fixed":
                carry["k"] = klast
            else:
                carry["k"] = f(u, t)
                stepper.nfev += 1
        return u
    assert kind == "tsit5"
    tend = t0 + dt
    t = t0
    saved = (stepper.dt_next, stepper.qold)   # a failed step leaves the controller as it was (like the engine's StepGuard)
    k1 = first_stage(t)  # FSAL invalidated by set_u! (carry is None) or kept (hybrid Z integrator)
    if stepper.dt_next is None:
        # Hairer-Norsett-Wanner initial step heuristic (order 5)
        sk = stepper.abstol + np.abs(u) * stepper.reltol
        d0 = math.sqrt(float(np.mean((u / sk) ** 2)))
        d1 = math.sqrt(float(np.mean((k1 / sk) ** 2)))
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        h0 = min(h0, dt)
        k1b = f(u + h0 * k1, t + h0)
        stepper.nfev += 1
        d2 = math.sqrt(float(np.mean(((k1b - k1) / sk) ** 2))) / h0
        h1 = max(1e-6, h0 * 1e-3) if max(d1, d2) <= 1e-15 else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
        stepper.dt_next = min(100.0 * h0, h1, dt)
    h = stepper.dt_next
    beta1, beta2, gamma, qmin, qmax = 7.0 / 50.0, 2.0 / 25.0, 0.9, 0.2, 10.0
    iters = 0
    while (tend - t) > 1e-14 * max(1.0, abs(tend)):
        iters += 1
        if iters > stepper.maxiters:
            stepper.dt_next, stepper.qold = saved
            raise MaxItersError("adaptive Tsit5 reached maxiters before the end of the step")
        h = min(h, tend - t)
        unew, ks = _tsit5_stages(f, u, t, h, k1)
        stepper.nfev += 6
        err = h * sum(b * k for b, k in zip(TSIT5_BTILDE, ks))
        sk = stepper.abstol + np.maximum(np.abs(u), np.abs(unew)) * stepper.reltol
        EEst = math.sqrt(float(np.mean((err / sk) ** 2)))
        if EEst <= 1.0:
            q11 = max(EEst, 1e-30) ** beta1
            q = q11 / (stepper.qold ** beta2)
            q = max(1.0 / qmax, min(1.0 / qmin, q / gamma))
            stepper.qold = max(EEst, 1e-4)
            t = t + h
            u = unew
            k1 = ks[6]
            stepper.naccept += 1
            hprop = h / q
            stepper.dt_next = hprop
            h = hprop
        else:
            q11 = EEst ** beta1
            q = min(1.0 / qmin, q11 / gamma)
            h = h / q
            stepper.nreject += 1
    if carry is not None:
        carry["k"] = k1
    return u


# ----------------------------------------------------------------------------------------------
# algorithms
# ----------------------------------------------------------------------------------------------


class PrimalLieTrotter:  # projector_splitting.jl:1
    pass


class DualLieTrotter:  # projector_splitting.jl:2
    pass


class Strang:  # projector_splitting.jl:3
    pass


def _stepper(x):
    return copy.deepcopy(x) if x is not None else SubStepper()


@dataclass
class ProjectorSplitting:  # projector_splitting.jl:31-41
    order: object = field(default_factory=PrimalLieTrotter)
    S_rhs: Optional[Callable] = None
    L_rhs: Optional[Callable] = None
    K_rhs: Optional[Callable] = None
    S_alg: Optional[SubStepper] = None
    L_alg: Optional[SubStepper] = None
    K_alg: Optional[SubStepper] = None


@dataclass
class UnconventionalAlgorithm:  # unconventional.jl:13-21
    S_rhs: Optional[Callable] = None
    L_rhs: Optional[Callable] = None
    K_rhs: Optional[Callable] = None
    S_alg: Optional[SubStepper] = None
    L_alg: Optional[SubStepper] = None
    K_alg: Optional[SubStepper] = None


@dataclass
class RankAdaptiveUnconventionalAlgorithm:  # rank_adaptive_unconventional.jl:15-23
    tol: float = 1e-8
    rmax: int = 2 ** 62
    S_rhs: Optional[Callable] = None
    L_rhs: Optional[Callable] = None
    K_rhs: Optional[Callable] = None
    S_alg: Optional[SubStepper] = None
    L_alg: Optional[SubStepper] = None
    K_alg: Optional[SubStepper] = None


@dataclass
class GreedyIntegrator:  # greedy_integrator.jl:16-22
    Z_alg: Optional[SubStepper] = None  # sub-stepper of the hybrid Z-flow (default Tsit5(), :19)
    fsal_carry: bool = True             # keep the Z integrator's cached first stage across outer steps (see ode_advance)


def _qr(A):
    """qr!(A); Matrix(Q), R  (LAPACK Householder thin QR; projector_splitting.jl:137-138 etc.)."""
    return np.linalg.qr(A, mode="reduced")


class _Cache:
    pass


def _alg_cache(prob, alg, u, dt, t0):
    """alg_cache for Data and DE problems (projector_splitting.jl:43-105, unconventional.jl:39-107,
    rank_adaptive_unconventional.jl:47-131).  The sub-"integrators" become closures that advance the
    K/S/L quantity by one outer step: for data problems `u + sign*left'*dy*right`
    (data_integrator.jl:15), for DE problems an explicit RK flow of the projected right-hand side."""
    c = _Cache()
    c.is_data = isinstance(prob, MatrixDataProblem)
    if isinstance(prob, MatrixHybridProblem):  # greedy_integrator.jl:41-47
        assert isinstance(alg, GreedyIntegrator), "MethodError: MatrixHybridProblem is solved by the GreedyIntegrator"
        c.feed = _DataFeed(prob.y, t0)
        c.f = prob.f
        c.Z_alg = _stepper(alg.Z_alg)
        c.Z_carry = {} if alg.fsal_carry else None
        return c
    if c.is_data:
        c.feed = _DataFeed(prob.y, t0)
        c.dy = None
    else:
        f = prob.f
        s_sign = -1.0 if isinstance(alg, ProjectorSplitting) else 1.0  # minus only in KSL (projector_splitting.jl:64)
        # default projected right-hand sides (projector_splitting.jl:52-80; unconventional.jl:51-79;
        # rank_adaptive_unconventional.jl:59-86).  Dense evaluation == Matrix(f(lowrank)*V) mathematically.
        c.K_rhs = alg.K_rhs or (lambda K, V, t: f(K @ V.T, t) @ V)
        c.L_rhs = alg.L_rhs or (lambda L, U, t: f(U @ L.T, t).T @ U)
        c.S_rhs = alg.S_rhs or (lambda S, UV, t: s_sign * (UV[0].T @ f(UV[0] @ S @ UV[1].T, t) @ UV[1]))
        c.K_alg, c.L_alg, c.S_alg = _stepper(alg.K_alg), _stepper(alg.L_alg), _stepper(alg.S_alg)
    if isinstance(alg, RankAdaptiveUnconventionalAlgorithm):
        c.r = u.rank
        c.tol = alg.tol
        c.r_max = alg.rmax
    return c


def _K_flow(c, K0, V, t, dt):
    if c.is_data:
        return K0 + data_contract(c.dy, None, V, 1)          # KIntegrator = (dy, US, I, u.V, +1)
    return ode_advance(c.K_alg, lambda K, tt: c.K_rhs(K, V, tt), K0, t, dt)


def _L_flow(c, L0, U, t, dt):
    if c.is_data:
        return L0 + data_contract(c.dy.T, None, U, 1)        # LIntegrator = (dy', VS, I, u.U, +1)
    return ode_advance(c.L_alg, lambda L, tt: c.L_rhs(L, U, tt), L0, t, dt)


def _S_flow(c, S0, U, V, t, dt, sign):
    if c.is_data:
        return S0 + data_contract(c.dy, U, V, sign)          # SIntegrator = (dy, ., u.U, u.V, -1|+1)
    return ode_advance(c.S_alg, lambda S, tt: c.S_rhs(S, (U, V), tt), S0, t, dt)


def primal_LT_step(u, c, t, dt, fetch=True):
    """projector_splitting.jl:117-152 (K -> S -> L)."""
    if c.is_data and fetch:
        c.dy = c.feed.advance(t, dt)
    K = _K_flow(c, u.U @ u.S, u.V, t, dt)                    # :133-136
    Q, R = _qr(K)                                            # :137
    u.U[...] = Q                                             # :138
    St = _S_flow(c, R, u.U, u.V, t, dt, -1)                  # :141-142 (S-step sees the NEW U, old V)
    L = _L_flow(c, u.V @ St.T, u.U, t, dt)                   # :145-148
    Q, R = _qr(L)                                            # :149
    u.V[...] = Q                                             # :150
    u.S[...] = R.T                                           # :151


def dual_LT_step(u, c, t, dt, fetch=True):
    """projector_splitting.jl:154-189 (L -> S -> K)."""
    if c.is_data and fetch:
        c.dy = c.feed.advance(t, dt)
    L = _L_flow(c, u.V @ u.S.T, u.U, t, dt)                  # :170-173
    Q, R = _qr(L)
    u.V[...] = Q                                             # :175
    St = _S_flow(c, R.T.copy(), u.U, u.V, t, dt, -1)         # :178-179
    K = _K_flow(c, u.U @ St, u.V, t, dt)                     # :182-185
    Q, R = _qr(K)
    u.U[...] = Q                                             # :187
    u.S[...] = R                                             # :188


def unconventional_step(u, c, t, dt):
    """unconventional.jl:121-157."""
    if c.is_data:
        c.dy = c.feed.advance(t, dt)
    K = _K_flow(c, u.U @ u.S, u.V, t, dt)                    # :137-140
    QK, _ = _qr(K)
    M = QK.T @ u.U                                           # :142
    L = _L_flow(c, u.V @ u.S.T, u.U, t, dt)                  # :145-148 (old U0: u.U not yet overwritten)
    QL, _ = _qr(L)
    N = QL.T @ u.V                                           # :150
    u.V[...] = QL                                            # :151
    u.U[...] = QK                                            # :152
    u.S[...] = _S_flow(c, M @ u.S @ N.T, u.U, u.V, t, dt, +1)  # :154-156


def rankadaptive_unconventional_step(u, c, t, dt):
    """rank_adaptive_unconventional.jl:182-233.  Returns (u_new | None, rank_adjusted)."""
    if c.is_data:
        c.dy = c.feed.advance(t, dt)
    r = c.r
    K = _K_flow(c, u.U @ u.S, u.V, t, dt)                    # :198-201
    Uhat, _ = _qr(np.hstack([K, u.U]))                       # :202-205
    M = Uhat.T @ u.U                                         # :206
    L = _L_flow(c, u.V @ u.S.T, u.U, t, dt)                  # :209-212
    Vhat, _ = _qr(np.hstack([L, u.V]))                       # :213-216
    N = Vhat.T @ u.V                                         # :217
    Shat = _S_flow(c, M @ u.S @ N.T, Uhat, Vhat, t, dt, +1)  # :219-220
    P, sig, Qt = np.linalg.svd(Shat)                         # :222
    r_new = min(c.r_max, truncate_to_tolerance(sig, c.tol))  # :223
    Unew = Uhat @ P[:, :r_new]
    Snew = np.diag(sig[:r_new])
    Vnew = Vhat @ Qt[:r_new, :].T
    if r_new == r:                                           # :224-228
        u.U[...] = Unew
        u.S[...] = Snew
        u.V[...] = Vnew
        return None, False
    return SVDLikeRepresentation(Unew, Snew, Vnew), True     # :230-231


def greedy_step(u, c, t, dt):
    """greedy_integrator.jl:94-104 (SVDLike, MatrixDataProblem): re-projection on the full snapshot X."""
    X = update_data(c.feed.y, t, dt)
    XV = X @ u.V
    XU = X.T @ u.U
    u.U[...] = _qr(XV)[0]
    u.V[...] = _qr(XU)[0]
    u.S[...] = u.U.T @ (X @ u.V)


def _polar_factor(XZ):
    """`Q, _, P = svd(XZ); mul!(U, Q, P')` (greedy_integrator.jl:79-80, 89-90): the orthogonal polar factor."""
    Q, _, Pt = np.linalg.svd(XZ, full_matrices=False)
    return Q @ Pt


def greedy_step_two_factor(u, c, t, dt):
    """greedy_integrator.jl:84-92 (TwoFactor, MatrixDataProblem): Z = X'U, then U = polar factor of X*Z."""
    X = update_data(c.feed.y, t, dt)
    u.Z[...] = X.T @ u.U
    u.U[...] = _polar_factor(X @ u.Z)


def greedy_step_hybrid(u, c, t, dt):
    """greedy_integrator.jl:72-82 (MatrixHybridProblem): Z advanced by its own ODE dZ/dt = f(Z, U, t) with
    the basis U as a (mutated in place) parameter, then U = polar factor of X(t+dt)*Z."""
    U_now = u.U  # the integrator's parameter p aliases u.U
    u.Z[...] = ode_advance(c.Z_alg, lambda Z, tt: c.f(Z, U_now, tt), u.Z, t, dt, carry=c.Z_carry)
    X = update_data(c.feed.y, t, dt)
    u.U[...] = _polar_factor(X @ u.Z)


def normal_component(U, Z, dY, C=None, tol=1e-8):
    """utils.jl:2-20: (I - U U') dY (I - Z pinv(C, atol=tol) Z'), C = Z'Z unless given.
    pinv(C; atol) (LinearAlgebra): singular values <= atol are dropped (rtol = 0 when atol > 0)."""
    if C is None:
        C = Z.T @ Z
    P, s, Qt = np.linalg.svd(np.asarray(C, dtype=np.float64))
    inv = np.where(s > tol, 1.0 / np.where(s > tol, s, 1.0), 0.0)
    Cp = (Qt.T * inv) @ P.T
    left = dY - U @ (U.T @ dY)
    return left - (left @ Z) @ Cp @ Z.T


def init(prob, alg, dt) -> DLRIntegrator:
    """projector_splitting.jl:107-115 | unconventional.jl:109-119 | rank_adaptive_unconventional.jl:94-104."""
    t0, tf = prob.tspan
    assert tf > t0, "Integration in reverse time direction is not supported"
    u = prob.u0.copy()
    sol = init_sol(dt, t0, tf, prob.u0)
    cache = _alg_cache(prob, alg, u, dt, t0)
    sol.Y[0] = prob.u0.copy()
    return DLRIntegrator(u, t0, dt, sol, alg, cache, type(prob), 0)


def step(integ: DLRIntegrator, alg=None, dt=None):
    """step!(integrator, alg, dt): projector_splitting.jl:191-211, unconventional.jl:159-164,
    rank_adaptive_unconventional.jl:171-180, greedy_integrator.jl:106-111."""
    alg = integ.alg if alg is None else alg
    dt = integ.dt if dt is None else dt
    u, t, c = integ.u, integ.t, integ.cache
    if isinstance(alg, ProjectorSplitting):
        if isinstance(alg.order, PrimalLieTrotter):
            primal_LT_step(u, c, t, dt)
        elif isinstance(alg.order, DualLieTrotter):
            dual_LT_step(u, c, t, dt)
        else:  # Strang: primal(dt/2) at t, dual(dt/2) at t+dt/2, each half fetching its own increment
            primal_LT_step(u, c, t, dt / 2)
            dual_LT_step(u, c, t + dt / 2, dt / 2)
    elif isinstance(alg, UnconventionalAlgorithm):
        unconventional_step(u, c, t, dt)
    elif isinstance(alg, RankAdaptiveUnconventionalAlgorithm):
        u_new, adjusted = rankadaptive_unconventional_step(u, c, t, dt)
        if adjusted:
            integ.u = u_new
            c.r = u_new.rank  # alg_recache (rank_adaptive_unconventional.jl:133-169): buffers re-sized for r_new
            if not c.is_data:  # ... and the ODE integrators are re-`init`ed (:150-164): fresh controller state
                for st in (c.K_alg, c.L_alg, c.S_alg):
                    st.dt_next, st.qold = None, 1e-4
    elif isinstance(alg, GreedyIntegrator):  # dispatch of greedy_step! on (typeof(u), probType), greedy_integrator.jl:72-104
        if integ.probType is MatrixHybridProblem:
            greedy_step_hybrid(u, c, t, dt)
        elif isinstance(u, TwoFactorRepresentation):
            greedy_step_two_factor(u, c, t, dt)
        else:
            greedy_step(u, c, t, dt)
    else:
        raise TypeError(f"MethodError: no step! for {type(alg).__name__}")
    integ.t += dt
    integ.iter += 1
