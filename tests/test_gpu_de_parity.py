"""GPU parity of the MatrixDEProblem steps (device-evaluable right-hand sides + explicit RK sub-steppers) against the
CPU oracle, through the C ABI.  Fixed-step sub-steppers are unambiguous; the adaptive Tsit5 shares one written spec
 between oracle and engine.  Bar: rel. Frobenius error of U·S·Vᵀ <= 1e-10 per step."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import dlra_oracle as O
from tests.problems import rel_fro, skew_pair

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def lri():
    import torch
    assert torch.cuda.is_available()
    import lowrankintegrators.jl_b200 as lri
    return lri


def dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)).cuda().t()


def csr_dev(M):
    import torch
    M = sp.csr_matrix(M)
    return (torch.from_numpy(M.indptr.astype(np.int64)).cuda(), torch.from_numpy(M.indices.astype(np.int32)).cuda(),
            torch.from_numpy(M.data.astype(np.float64)).cuda(), M.shape)


def periodic_ops(n, nu=0.005, length=np.pi):
    """test/data_agnostic_approximation.jl:6-29: periodic centred Laplacian (times viscosity) and gradient."""
    dx = length / n
    i = np.arange(n)
    lap = sp.csr_matrix((np.r_[np.full(n, nu / dx ** 2), np.full(n, -2 * nu / dx ** 2), np.full(n, nu / dx ** 2)],
                         (np.r_[i, i, i], np.r_[(i - 1) % n, i, (i + 1) % n])), shape=(n, n))
    grad = sp.csr_matrix((np.r_[np.full(n, -0.5 / dx), np.full(n, 0.5 / dx)], (np.r_[i, i], np.r_[(i - 1) % n, (i + 1) % n])), shape=(n, n))
    return lap, grad


def algs(lri, sub, osub):
    kw = dict(K_alg=sub, L_alg=sub, S_alg=sub)
    okw = dict(K_alg=osub, L_alg=osub, S_alg=osub)
    return {
        "bug": (lri.UnconventionalAlgorithm(**kw), O.UnconventionalAlgorithm(**okw)),
        "ksl_primal": (lri.ProjectorSplitting(lri.PrimalLieTrotter(), **kw), O.ProjectorSplitting(O.PrimalLieTrotter(), **okw)),
        "ksl_dual": (lri.ProjectorSplitting(lri.DualLieTrotter(), **kw), O.ProjectorSplitting(O.DualLieTrotter(), **okw)),
        "ksl_strang": (lri.ProjectorSplitting(lri.Strang(), **kw), O.ProjectorSplitting(O.Strang(), **okw)),
        "rabug": (lri.RankAdaptiveUnconventionalAlgorithm(1e-7, rmax=12, **kw), O.RankAdaptiveUnconventionalAlgorithm(1e-7, rmax=12, **okw)),
    }


def run_both(lri, grhs, of, X0, galg, oalg, dt, nsteps, resync=True):
    gu0 = lri.SVDLikeRepresentation(X0.U, X0.S, X0.V)
    gint = lri.init(lri.MatrixDEProblem(grhs, gu0, (0.0, dt * nsteps)), galg, dt)
    oint = O.init(O.MatrixDEProblem(of, X0, (0.0, dt * nsteps)), oalg, dt)
    errs = []
    for k in range(nsteps):
        if resync:
            gint.cache.set_factors(oint.u.U, oint.u.S, oint.u.V)
        O.step(oint)
        lri.step(gint)
        gu, ou = gint.u, oint.u
        assert gu.rank == ou.rank, f"rank mismatch at step {k}: {gu.rank} vs {ou.rank}"
        errs.append(rel_fro(gu.full(), ou.full()))
    return errs


SUBS = [("rk4", 2), ("tsit5_fixed", 1), ("euler", 4), ("tsit5", 1)]


@pytest.mark.parametrize("name", ["bug", "ksl_primal", "ksl_dual", "ksl_strang", "rabug"])
@pytest.mark.parametrize("sub", SUBS)
def test_linear_generic_matrix_de(lri, name, sub):
    # DE form of examples/generic_matrix.jl:16: F(X) = W1·X + X + X·W2, dense operators
    N, r = 96, 5
    W1, W2 = (0.2 * W for W in skew_pair(N, seed=4))
    D = np.diag(2.0 ** -np.arange(1, N + 1))
    X0 = O.truncated_svd(D + 1e-3 * np.random.default_rng(0).standard_normal((N, N)), r)
    of = lambda X, t: W1 @ X + X + X @ W2
    grhs = lri.LinearRHS(A=dev(W1 + np.eye(N)), B=dev(W2.T))
    galg, oalg = algs(lri, lri.SubStepper(*sub), O.SubStepper(*sub))[name]
    errs = run_both(lri, grhs, of, X0, galg, oalg, 0.01, 3)
    assert max(errs) <= TOL, errs


@pytest.mark.parametrize("name", ["bug", "ksl_primal", "rabug"])
def test_lyapunov_with_forcing_csr(lri, name):
    # config 4 style: F(X) = A·X + X·Bᵀ + G·Hᵀ with stencil operators (CSR) and a low-rank source that grows the rank
    n, m, r = 384, 256, 4
    rng = np.random.default_rng(3)
    A = sum(periodic_ops(n, nu=0.02))
    B = sum(periodic_ops(m, nu=0.03))
    G, H = rng.standard_normal((n, 6)), rng.standard_normal((m, 6))
    G, H = np.linalg.qr(G)[0], np.linalg.qr(H)[0] * (2.0 ** -np.arange(6))
    X0 = O.truncated_svd(rng.standard_normal((n, r)) @ np.diag(2.0 ** -np.arange(r)) @ rng.standard_normal((r, m)), r)
    Ad, Bd = A.toarray(), B.toarray()
    of = lambda X, t: Ad @ X + X @ Bd.T + G @ H.T
    grhs = lri.LinearRHS(A=csr_dev(A), B=csr_dev(B), G=dev(G), H=dev(H))
    sub = ("rk4", 4)
    galg, oalg = algs(lri, lri.SubStepper(*sub), O.SubStepper(*sub))[name]
    errs = run_both(lri, grhs, of, X0, galg, oalg, 0.005, 4, resync=(name != "rabug"))
    assert max(errs) <= (1e-9 if name == "rabug" else TOL), errs


@pytest.mark.parametrize("name", ["bug", "ksl_primal", "ksl_dual", "ksl_strang", "rabug"])
@pytest.mark.parametrize("sub", [("rk4", 2), ("tsit5", 1)])
def test_burgers_uncertainty(lri, name, sub):
    # test/data_agnostic_approximation.jl:4-49 (Burgers UQ, F(ρ) = Δρ − (∇ρ).*ρ) at n=256, m=12², rank 5
    n, mm, r = 256, 12, 5
    x = (np.arange(n) + 0.5) * np.pi / n
    lap, grad = periodic_ops(n)
    xi = [(a, b) for b in np.linspace(-1, 1, mm) for a in np.linspace(-1, 1, mm)]
    ub = 0.5 * (np.exp(np.cos(x)) - 1.5) * np.sin(x + 2 * np.pi * 0.37)
    rho0 = np.stack([ub + 0.5 * a * np.sin(2 * np.pi * x) + 0.5 * b * np.sin(3 * np.pi * x) + 0.1 * a * b * np.cos(4 * x)
                     + 0.05 * (a * a - b) * np.sin(5 * x) for a, b in xi], axis=1)
    X0 = O.truncated_svd(rho0, r)     # exactly rank 5 => well-conditioned S0
    Ld, Gd = lap.toarray(), grad.toarray()
    of = lambda X, t: Ld @ X - (Gd @ X) * X
    grhs = lri.BurgersRHS(csr_dev(lap), csr_dev(grad))
    galg, oalg = algs(lri, lri.SubStepper(*sub), O.SubStepper(*sub))[name]
    if name == "rabug":
        galg, oalg = (lri.RankAdaptiveUnconventionalAlgorithm(1e-4, rmax=10, K_alg=galg.K_alg, L_alg=galg.L_alg, S_alg=galg.S_alg),
                      O.RankAdaptiveUnconventionalAlgorithm(1e-4, rmax=10, K_alg=oalg.K_alg, L_alg=oalg.L_alg, S_alg=oalg.S_alg))
    errs = run_both(lri, grhs, of, X0, galg, oalg, 0.01, 3)
    assert max(errs) <= TOL, errs


def test_de_solve_matches_oracle_trajectory(lri):
    N, r = 64, 4
    W1, W2 = (0.1 * W for W in skew_pair(N, seed=9))
    X0 = O.truncated_svd(np.diag(2.0 ** -np.arange(1, N + 1)), r)
    of = lambda X, t: W1 @ X + X + X @ W2
    grhs = lri.LinearRHS(A=dev(W1 + np.eye(N)), B=dev(W2.T))
    sol = lri.solve(lri.MatrixDEProblem(grhs, lri.SVDLikeRepresentation(X0.U, X0.S, X0.V), (0.0, 0.2)),
                    lri.ProjectorSplitting(lri.PrimalLieTrotter()), 0.01)
    osol = O.solve(O.MatrixDEProblem(of, X0, (0.0, 0.2)), O.ProjectorSplitting(O.PrimalLieTrotter()), 0.01)
    assert len(sol.Y) == len(osol.Y) and abs(sol.t[-1] - osol.t[-1]) < 1e-12
    assert rel_fro(sol.Y[-1].full(), osol.Y[-1].full()) <= 1e-9


def test_reference_burgers_test_config(lri):
    """test/data_agnostic_approximation.jl as written: n = 1000, m = 20^2, r = 5, dt = 1e-2, the five solvers with their DEFAULT
    sub-integrators (adaptive Tsit5, abstol 1e-6, reltol 1e-3).  Three steps each against the oracle."""
    import torch
    n, mm, r, dt = 1000, 20, 5, 1e-2
    lap, grad = periodic_ops(n)
    x = (np.arange(n) + 0.5) * (np.pi / n)
    ub = 0.5 * (np.exp(np.cos(x)) - 1.5) * np.sin(x + 2 * np.pi * 0.37)
    xi = [(a, b) for a in np.linspace(-1, 1, mm) for b in np.linspace(-1, 1, mm)]
    rho0 = np.stack([ub + 0.5 * a * np.sin(2 * np.pi * x) + 0.5 * b * np.sin(3 * np.pi * x) for a, b in xi], axis=1)
    X0 = O.truncated_svd(rho0, r)
    of = lambda rho, t: lap @ rho - (grad @ rho) * rho
    grhs = lri.BurgersRHS(csr_dev(lap), csr_dev(grad))
    pairs = {
        "bug": (lri.UnconventionalAlgorithm(), O.UnconventionalAlgorithm()),
        "ksl_primal": (lri.ProjectorSplitting(lri.PrimalLieTrotter()), O.ProjectorSplitting(O.PrimalLieTrotter())),
        "ksl_dual": (lri.ProjectorSplitting(lri.DualLieTrotter()), O.ProjectorSplitting(O.DualLieTrotter())),
        "ksl_strang": (lri.ProjectorSplitting(lri.Strang()), O.ProjectorSplitting(O.Strang())),
        "rabug": (lri.RankAdaptiveUnconventionalAlgorithm(1e-4, rmax=10), O.RankAdaptiveUnconventionalAlgorithm(1e-4, rmax=10)),
    }
    for name, (galg, oalg) in pairs.items():
        errs = run_both(lri, grhs, of, X0, galg, oalg, dt, 3, resync=True)
        assert max(errs) <= 1e-9, (name, errs)   # adaptive sub-steppers: accept/reject sequences must coincide


def test_stiff_projected_flow_hits_maxiters_like_the_oracle(lri):
    """BASELINE configs[2] at its stated dt = 1e-2 is outside the explicit sub-stepper's reach once the grid is fine: the
    discrete Laplacian has |lambda|max = 4 nu / dx^2 (1.36e5 at n = 8192), Tsit5 is stable for h*|lambda| < ~3.5, so one
    outer step needs > dt*|lambda|/3.5 sub-steps per flow.  OrdinaryDiffEq stops at maxiters (default 1e5) with retcode
    MaxIters; engine and oracle must do the same thing at the same point, and the engine must stay usable afterwards."""
    n, mm, r, dt = 2048, 8, 8, 1e-2
    lap, grad = periodic_ops(n)
    x = (np.arange(n) + 0.5) * (np.pi / n)
    ub = 0.5 * (np.exp(np.cos(x)) - 1.5) * np.sin(x + 2 * np.pi * 0.37)
    xi = [(a, b) for a in np.linspace(-1, 1, mm) for b in np.linspace(-1, 1, mm)]
    rho0 = np.stack([ub + 0.5 * a * np.sin(2 * np.pi * x) + 0.5 * b * np.sin(3 * np.pi * x) for a, b in xi], axis=1)
    X0 = O.truncated_svd(rho0, r)
    of = lambda rho, t: lap @ rho - (grad @ rho) * rho
    grhs = lri.BurgersRHS(csr_dev(lap), csr_dev(grad))
    lam = 4 * 0.005 / (np.pi / n) ** 2
    need = dt * lam / 3.5            # ~24 sub-steps per flow at n = 2048
    for maxiters, expect_fail in ((int(need // 2), True), (4000, False)):
        gsub = lambda: lri.SubStepper(maxiters=maxiters)
        osub = lambda: O.SubStepper(maxiters=maxiters)
        galg = lri.ProjectorSplitting(lri.PrimalLieTrotter(), K_alg=gsub(), S_alg=gsub(), L_alg=gsub())
        oalg = O.ProjectorSplitting(O.PrimalLieTrotter(), K_alg=osub(), S_alg=osub(), L_alg=osub())
        gint = lri.init(lri.MatrixDEProblem(grhs, lri.SVDLikeRepresentation(X0.U, X0.S, X0.V), (0.0, 1.0)), galg, dt)
        oint = O.init(O.MatrixDEProblem(of, X0, (0.0, 1.0)), oalg, dt)
        if expect_fail:
            with pytest.raises(O.MaxItersError):
                O.step(oint)
            with pytest.raises(lri._lib.DLRAError) as ei:
                lri.step(gint)
            assert ei.value.code == lri._lib.EMAXITERS
            # the failed step left the factors untouched and the handle usable
            U, S, V = gint.cache.get_factors()
            assert rel_fro(U @ S @ V.T, X0.full()) < 1e-14
            for f in (lri._lib.FLOW_K, lri._lib.FLOW_S, lri._lib.FLOW_L):
                gint.cache.set_substepper(f, lri._lib.ODE_TSIT5, 1, 1e-6, 1e-3, maxiters=4000)
            gint.cache.step_ksl(lri._lib.KSL_PRIMAL, 0.0, dt)
            for s in (oint.cache.K_alg, oint.cache.S_alg, oint.cache.L_alg):
                s.maxiters = 4000
            O.step(oint)
            U, S, V = gint.cache.get_factors()
            assert rel_fro(U @ S @ V.T, oint.u.full()) <= 1e-9
        else:
            O.step(oint)
            lri.step(gint)
            assert rel_fro(gint.u.full(), oint.u.full()) <= 1e-9
        gint.cache.close()
