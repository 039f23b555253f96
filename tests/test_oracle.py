"""Pins for the CPU oracle (the reference holds no golden vectors):
(1) restatement of the reference's only executed assertion (test/data_driven_approximation.jl:30),
(2) exactness on rank-r streams, (3) analytic best rank-r error of examples/generic_matrix.jl."""
import numpy as np
import pytest

from oracle import dlra_oracle as O
from tests.problems import generic_matrix_stream, lowrank_stream, rel_fro

ALGS = {
    "bug": lambda: O.UnconventionalAlgorithm(),
    "ksl_primal": lambda: O.ProjectorSplitting(O.PrimalLieTrotter()),
    "ksl_dual": lambda: O.ProjectorSplitting(O.DualLieTrotter()),
    "rabug": lambda: O.RankAdaptiveUnconventionalAlgorithm(1e-8, rmax=20),
    "greedy": lambda: O.GreedyIntegrator(),
}


@pytest.mark.parametrize("name", list(ALGS))
def test_reference_data_compression_selfconsistency(name):
    # test/data_driven_approximation.jl:2-30 with a seeded RNG
    Y = generic_matrix_stream(100, seed=0)
    X0 = O.truncated_svd(Y(0.0), tol=1e-4)
    prob = O.MatrixDataProblem(Y, X0, (0.0, 1.0))
    data = [Y(t) for t in np.arange(0, 1.0 + 1e-12, 0.01)]
    assert len(data) == 101
    dprob = O.MatrixDataProblem(data, X0)
    assert dprob.tspan == (1, 101)
    sol = O.solve(prob, ALGS[name](), 1e-2)
    dsol = O.solve(dprob, ALGS[name]())
    assert len(dsol.Y) == 101 and dsol.t[-1] == 101
    A, B = sol.Y[-1].full(), dsol.Y[-1].full()
    assert np.allclose(A, B, rtol=np.sqrt(np.finfo(float).eps), atol=1e-12)


@pytest.mark.parametrize("name", ["bug", "ksl_primal", "ksl_dual"])
def test_exactness_on_rank_r_stream(name):
    # README.md refs [1],[2]: rank A(t) <= r  =>  KSL and BUG reproduce A(t_k) to round-off
    A = lowrank_stream(300, 200, 6, seed=1)
    snaps = [A(0.05 * k) for k in range(21)]
    X0 = O.truncated_svd(snaps[0], 6)
    sol = O.solve(O.MatrixDataProblem(snaps, X0), ALGS[name]())
    assert rel_fro(sol.Y[-1].full(), snaps[-1]) < 1e-13


def test_strang_on_snapshot_vector_is_method_error():
    A = lowrank_stream(30, 20, 3, seed=2)
    snaps = [A(0.1 * k) for k in range(4)]
    X0 = O.truncated_svd(snaps[0], 3)
    with pytest.raises(AssertionError):
        O.solve(O.MatrixDataProblem(snaps, X0), O.ProjectorSplitting(O.Strang()))


@pytest.mark.parametrize("r,best", [(4, 0.0981), (8, 6.13e-3)])
def test_generic_matrix_error_approaches_best_rank_r(r, best):
    # examples/generic_matrix.jl:16-18,34: sigma_j(Y(1)) = e*2^-j  =>  best rank-r error e*2^-r/sqrt(3)
    Y = generic_matrix_stream(100, seed=0)
    assert abs(np.e * 2.0 ** -r / np.sqrt(3) - best) / best < 1e-2
    X0 = O.truncated_svd(Y(0.0), r)
    prob = O.MatrixDataProblem(Y, X0, (0.0, 1.0))
    for alg in (ALGS["bug"](), ALGS["ksl_primal"](), O.ProjectorSplitting(O.Strang())):
        sol = O.solve(prob, alg, 0.01)
        err = np.linalg.norm(sol.Y[-1].full() - Y(1.0))
        assert best * 0.999 <= err <= 2.5 * best


def test_truncate_to_tolerance():
    s = np.array([1.0, 1e-1, 1e-2, 1e-3])
    assert O.truncate_to_tolerance(s, 2e-3) == 3
    assert O.truncate_to_tolerance(s, 1e-3) == 3
    assert O.truncate_to_tolerance(s, 0.9e-3) == 4
    assert O.truncate_to_tolerance(s, 10.0) == 0


def test_rank_adaptive_grows_and_caps():
    A = lowrank_stream(120, 90, 12, seed=3)
    snaps = [A(0.2 * k) for k in range(8)]
    X0 = O.truncated_svd(snaps[0], 3)
    sol = O.solve(O.MatrixDataProblem(snaps, X0), O.RankAdaptiveUnconventionalAlgorithm(1e-6, rmax=10))
    ranks = [y.rank for y in sol.Y]
    assert ranks[0] == 3 and max(ranks) == 10 and all(b <= 2 * a for a, b in zip(ranks, ranks[1:]))


def test_de_problem_linear_matches_exact_flow():
    # F(X) = W1 X + X + X W2 has the exact solution of generic_matrix.jl; rank-r start stays close.
    from tests.problems import skew_pair
    from scipy.linalg import expm
    N, r = 60, 6
    W1, W2 = (0.05 * W for W in skew_pair(N, seed=4))  # mild rotation: BUG is first order in dt*||W||
    D = np.diag(2.0 ** -np.arange(1, N + 1))
    X0 = O.truncated_svd(D, r)
    f = lambda X, t: W1 @ X + X + X @ W2
    exact = expm(0.5 * W1) @ (np.exp(0.5) * X0.full()) @ expm(0.5 * W2)  # rank-r data stays rank r
    for alg in (O.UnconventionalAlgorithm(), O.ProjectorSplitting(O.PrimalLieTrotter()),
                O.ProjectorSplitting(O.Strang()), O.RankAdaptiveUnconventionalAlgorithm(1e-10, rmax=12)):
        sol = O.solve(O.MatrixDEProblem(f, X0, (0.0, 0.5)), alg, 0.05)
        assert rel_fro(sol.Y[-1].full(), exact) < 5e-3
    rk4 = O.SubStepper("rk4", nsub=4)
    errs = []
    for dt in (0.05, 0.025):
        sol = O.solve(O.MatrixDEProblem(f, X0, (0.0, 0.5)),
                      O.UnconventionalAlgorithm(K_alg=rk4, L_alg=rk4, S_alg=rk4), dt)
        errs.append(rel_fro(sol.Y[-1].full(), exact))
    assert 1.6 < errs[0] / errs[1] < 2.4  # first-order convergence of fixed-rank BUG
    sol = O.solve(O.MatrixDEProblem(f, X0, (0.0, 0.5)),
                  O.ProjectorSplitting(O.PrimalLieTrotter(), K_alg=rk4, L_alg=rk4, S_alg=rk4), 0.05)
    assert rel_fro(sol.Y[-1].full(), exact) < 1e-8  # KSL is exact on rank-preserving linear flows


# ---- greedy integrator on u = U*Z' and the hybrid problem (greedy_integrator.jl:72-92, utils.jl:2-20) ----

def _linear_lowrank_flow(n=60, m=40, r=5, seed=0):
    from scipy.linalg import expm
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)); A = 0.3 * (A - A.T)
    B = rng.standard_normal((m, m)); B = 0.3 * (B - B.T)
    Y0 = (rng.standard_normal((n, r)) * 2.0 ** -np.arange(r)) @ rng.standard_normal((r, m))
    y = lambda t: expm(t * A) @ Y0 @ expm(t * B.T)
    F = lambda X, t: A @ X + X @ B.T
    U0 = np.linalg.svd(Y0, full_matrices=False)[0][:, :r]
    return y, F, O.TwoFactorRepresentation(U0, Y0.T @ U0)


@pytest.mark.parametrize("carry", [True, False])
def test_hybrid_greedy_tracks_rank_r_flow(carry):
    # y(t) has rank r and solves Y' = F(Y): the Galerkin coefficients Z' = F(UZ')'U together with the polar update of U
    # reproduce it up to the ODE tolerance, with or without the carried first stage
    y, F, u0 = _linear_lowrank_flow()
    prob = O.MatrixHybridProblem(y, lambda Z, U, t: F(U @ Z.T, t).T @ U, u0, (0.0, 0.5))
    sol = O.solve(prob, O.GreedyIntegrator(Z_alg=O.SubStepper("tsit5", abstol=1e-10, reltol=1e-8), fsal_carry=carry), 0.05)
    assert len(sol.Y) == 11 and isinstance(sol.Y[-1], O.TwoFactorRepresentation)
    assert rel_fro(sol.Y[-1].full(), y(0.5)) < 1e-6
    U = sol.Y[-1].U
    assert np.linalg.norm(U.T @ U - np.eye(U.shape[1])) < 1e-12


def test_two_factor_greedy_data_step_is_polar_update():
    y, _, u0 = _linear_lowrank_flow(seed=3)
    snaps = [y(0.05 * k) for k in range(4)]
    integ = O.init(O.MatrixDataProblem(snaps, u0), O.GreedyIntegrator(), 1)
    Uold = integ.u.U.copy()
    O.step(integ)
    X = snaps[1]
    assert np.allclose(integ.u.Z, X.T @ Uold, atol=1e-14)                    # mul!(Z, X', U)   (:87)
    G = X @ integ.u.Z
    Unew = integ.u.U
    assert np.linalg.norm(Unew.T @ Unew - np.eye(5)) < 1e-12
    H = Unew.T @ G                                                             # polar factor: U'(XZ) symmetric positive semidefinite
    assert np.allclose(H, H.T, atol=1e-12) and np.linalg.eigvalsh(0.5 * (H + H.T)).min() > -1e-12


def test_normal_component_properties():
    rng = np.random.default_rng(5)
    n, m, r = 50, 30, 4
    U = np.linalg.qr(rng.standard_normal((n, r)))[0]
    Z = rng.standard_normal((m, r))
    dY = rng.standard_normal((n, m))
    N = O.normal_component(U, Z, dY)
    assert np.linalg.norm(U.T @ N) < 1e-12 and np.linalg.norm(N @ Z) < 1e-12  # orthogonal to range(U) and to range(Z)
    T = U @ rng.standard_normal((r, m)) + rng.standard_normal((n, r)) @ Z.T   # tangent directions have no normal component
    assert np.linalg.norm(O.normal_component(U, Z, T)) < 1e-11
    # rank-deficient Z: pinv(C, atol) drops the null direction instead of blowing up
    Z2 = Z.copy(); Z2[:, 3] = Z2[:, 2]
    N2 = O.normal_component(U, Z2, dY)
    assert np.isfinite(N2).all() and np.linalg.norm(N2 @ Z2) < 1e-10


def test_reference_data_informed_burgers_reduced():
    # test/data_informed_approximation.jl at reduced size (n=128 instead of 1000, 6^2 samples instead of 10^2, rank 10, and
    # viscosity 0.02 instead of 0.005 so that the coarser grid resolves the front): <= 10 % error
    from tests.problems import burgers_truth
    n, mm, r, dt = 128, 6, 10, 1e-2
    xi = [(a, b) for b in np.linspace(-1, 1, mm) for a in np.linspace(-1, 1, mm)]
    t_grid = np.arange(0, 0.3 + 1e-12, dt)
    truth, F = burgers_truth(n, xi, t_grid, nu=0.02)
    U0 = O.truncated_svd(np.hstack([truth[0], truth[1]]), r).U               # :67-69
    u0 = O.TwoFactorRepresentation(U0, truth[0].T @ U0)
    y = lambda t: truth[min(int(np.floor(t / dt + 1e-9)), len(t_grid) - 1)]    # interpolate_data (:70-73)
    prob = O.MatrixHybridProblem(y, lambda Z, U, t: F(U @ Z.T).T @ U, u0, (0.0, 0.3))
    sol = O.solve(prob, O.GreedyIntegrator(), dt)
    assert rel_fro(sol.Y[-1].full(), truth[-1]) <= 0.1


# ---- sub-steppers: the restated Tsit5 tableau against its published properties (OrdinaryDiffEq itself is not available) ----

def test_tsit5_tableau_consistency():
    A, c, bt = O.TSIT5_A, O.TSIT5_C, O.TSIT5_BTILDE
    for s in range(1, 7):
        assert abs(sum(A[s]) - c[s]) < 1e-14                      # row sums = nodes
    b = np.array(list(A[6]) + [0.0])                               # FSAL: the 7th stage row is the 5th-order weight vector
    cc = np.array(c)
    Am = np.zeros((7, 7))
    for s in range(1, 7):
        Am[s, :len(A[s])] = A[s]
    # order conditions up to order 4 (all of them) and the quadrature conditions of order 5
    assert abs(b.sum() - 1) < 1e-14 and abs(b @ cc - 1 / 2) < 1e-14 and abs(b @ cc ** 2 - 1 / 3) < 1e-14
    assert abs(b @ (Am @ cc) - 1 / 6) < 1e-14 and abs(b @ cc ** 3 - 1 / 4) < 1e-14
    assert abs(b @ (cc * (Am @ cc)) - 1 / 8) < 1e-14 and abs(b @ (Am @ cc ** 2) - 1 / 12) < 1e-14
    assert abs(b @ (Am @ (Am @ cc)) - 1 / 24) < 1e-14 and abs(b @ cc ** 4 - 1 / 5) < 1e-13
    bt = np.array(bt)
    assert abs(bt.sum()) < 1e-14                                   # error weights: difference of two consistent methods
    bhat = b - bt                                                  # embedded 4th-order solution
    assert abs(bhat @ cc - 1 / 2) < 1e-14 and abs(bhat @ cc ** 2 - 1 / 3) < 1e-14 and abs(bhat @ cc ** 3 - 1 / 4) < 1e-13


@pytest.mark.parametrize("kind,order", [("euler", 1), ("rk4", 4), ("tsit5_fixed", 5)])
def test_fixed_substeppers_converge_with_their_order(kind, order):
    rng = np.random.default_rng(2)
    M = rng.standard_normal((4, 4)); M = 0.5 * (M - M.T) - 0.1 * np.eye(4)
    f = lambda u, t: M @ u + np.sin(3 * t) * np.ones((4, 1)) * 0.3 + 0.2 * u * u       # nonlinear, non-autonomous
    u0 = rng.standard_normal((4, 1))
    ref = O.ode_advance(O.SubStepper("tsit5_fixed", nsub=2000), f, u0, 0.0, 1.0)
    errs = [np.linalg.norm(O.ode_advance(O.SubStepper(kind, nsub=ns), f, u0, 0.0, 1.0) - ref) for ns in (20, 40)]
    assert abs(np.log2(errs[0] / errs[1]) - order) < 0.35, errs


def test_adaptive_tsit5_meets_tolerance_and_reaches_the_end_exactly():
    from scipy.integrate import solve_ivp
    rng = np.random.default_rng(3)
    M = rng.standard_normal((6, 6)); M = M - M.T
    f = lambda u, t: M @ u - 0.5 * u ** 3
    u0 = rng.standard_normal(6)
    exact = solve_ivp(lambda t, u: f(u, t), (0.0, 2.0), u0, rtol=1e-12, atol=1e-14, method="DOP853").y[:, -1]
    prev = None
    for tol in (1e-4, 1e-6, 1e-8):
        st = O.SubStepper("tsit5", abstol=tol * 1e-3, reltol=tol)
        u = u0.copy()
        for k in range(4):                                        # four outer steps: the controller state carries over
            u = O.ode_advance(st, f, u, 0.5 * k, 0.5)
        err = np.linalg.norm(u - exact) / np.linalg.norm(exact)
        assert err < 50 * tol, (tol, err)
        assert prev is None or st.naccept > prev                  # tighter tolerance -> more steps
        assert st.nreject <= st.naccept
        prev = st.naccept


def test_reference_markov_chain_example_reduced():
    # examples/markov_chain.jl at reduced size (N=64 instead of 600, Tf=0.4): BUG on the chemical master equation, written as the
    # sum of two-sided terms Σ_k A_k·P·B_kᵀ the engine evaluates (dlra_rhs_add_term), tracks the directly integrated solution
    from tests.problems import cme_operators
    N = 64
    terms = [(A.toarray(), B.toarray()) for A, B in cme_operators(N)]
    f = lambda P, t: sum(A @ P @ B.T for A, B in terms)
    xs = np.arange(1, N + 1)
    D = np.array([[0.03, 0.01], [0.01, 0.02]])
    P0 = np.array([[np.exp(-np.array([x - 20, y - 20]) @ D @ np.array([x - 20, y - 20])) for y in xs] for x in xs])
    P0 /= P0.sum()
    X0 = O.truncated_svd(P0, tol=1e-6)                                          # markov_chain.jl:86
    dt, Tf = 2e-2, 0.4
    sol = O.solve(O.MatrixDEProblem(f, X0, (0.0, Tf)), O.UnconventionalAlgorithm(), dt)
    P, h = P0.copy(), dt / 20
    for _ in range(int(round(Tf / h))):
        k1 = f(P, 0); k2 = f(P + 0.5 * h * k1, 0); k3 = f(P + 0.5 * h * k2, 0); k4 = f(P + h * k3, 0)
        P = P + (h / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
    assert abs(P.sum() - 1) < 1e-9                                              # probability is conserved (mass only leaves at the truncation boundary)
    assert rel_fro(sol.Y[-1].full(), P) < 1e-3
