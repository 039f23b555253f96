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
