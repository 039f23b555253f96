"""Golden fixtures (tests/golden/oracle_golden.npz, made by tests/golden/make_golden.py).
CPU: the oracle must still reproduce them.  GPU: the engine must match them through the C ABI without importing the oracle."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"))
DATA_ALGS = ["bug", "ksl_primal", "ksl_dual", "rabug", "greedy"]


def rel(X, Y):
    return np.linalg.norm(X - Y) / np.linalg.norm(Y)


@pytest.mark.parametrize("name", DATA_ALGS)
def test_oracle_reproduces_golden_data_problem(name):
    from oracle import dlra_oracle as O
    algs = {"bug": O.UnconventionalAlgorithm(), "ksl_primal": O.ProjectorSplitting(O.PrimalLieTrotter()),
            "ksl_dual": O.ProjectorSplitting(O.DualLieTrotter()), "rabug": O.RankAdaptiveUnconventionalAlgorithm(1e-2, rmax=8),
            "greedy": O.GreedyIntegrator()}
    X0 = O.SVDLikeRepresentation(G["data_U0"], G["data_S0"], G["data_V0"])
    sol = O.solve(O.MatrixDataProblem(list(G["data_snaps"]), X0), algs[name])
    assert [y.rank for y in sol.Y] == list(G[f"data_{name}_rank"])
    assert max(rel(y.full(), g) for y, g in zip(sol.Y, G[f"data_{name}_Y"])) < 1e-12


def _hybrid_inputs():
    snaps = G["hybrid_snaps"]
    y = lambda t: snaps[min(int(round(t / 0.01)), len(snaps) - 1)]
    U0 = G["de_U0"]
    Y0 = U0 @ G["de_S0"] @ G["de_V0"].T
    return y, U0, Y0.T @ U0


def test_oracle_reproduces_golden_two_factor_and_hybrid():
    from oracle import dlra_oracle as O
    Z0 = G["data_snaps"][0].T @ G["data_U0"]
    sol = O.solve(O.MatrixDataProblem(list(G["data_snaps"]), O.TwoFactorRepresentation(G["data_U0"], Z0)), O.GreedyIntegrator())
    assert max(rel(y.full(), g) for y, g in zip(sol.Y, G["data_greedy2_Y"])) < 1e-12
    W1, W2 = G["de_W1"], G["de_W2"]
    fz = lambda Z, U, t: (W1 @ (U @ Z.T) + U @ Z.T + (U @ Z.T) @ W2).T @ U
    y, U0, Zh = _hybrid_inputs()
    for carry in (True, False):
        alg = O.GreedyIntegrator(Z_alg=O.SubStepper("rk4", nsub=2), fsal_carry=carry)
        sol = O.solve(O.MatrixHybridProblem(y, fz, O.TwoFactorRepresentation(U0, Zh), (0.0, 0.04)), alg, 0.01)
        assert max(rel(s.full(), g) for s, g in zip(sol.Y, G[f"hybrid_carry{int(carry)}_Y"])) < 1e-12
        assert max(rel(s.Z, g) for s, g in zip(sol.Y, G[f"hybrid_carry{int(carry)}_Z"])) < 1e-11
    assert rel(G["hybrid_carry1_Z"][-1], G["hybrid_carry0_Z"][-1]) > 1e-6   # the fixtures do tell the two variants apart


@pytest.mark.gpu
def test_engine_matches_golden_two_factor_and_hybrid():
    import torch
    import lowrankintegrators.jl_b200 as lri
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(np.asarray(x).T)).cuda().t()
    Z0 = G["data_snaps"][0].T @ G["data_U0"]
    sol = lri.solve(lri.MatrixDataProblem(list(G["data_snaps"]), lri.TwoFactorRepresentation(G["data_U0"], Z0)), lri.GreedyIntegrator())
    assert max(rel(y.full(), g) for y, g in zip(sol.Y, G["data_greedy2_Y"])) <= 1e-10
    N = G["de_W1"].shape[0]
    y, U0, Zh = _hybrid_inputs()
    for carry in (True, False):
        rhs = lri.LinearRHS(A=dev(G["de_W1"] + np.eye(N)), B=dev(G["de_W2"].T))
        alg = lri.GreedyIntegrator(Z_alg=lri.SubStepper("rk4", nsub=2), fsal_carry=carry)
        sol = lri.solve(lri.MatrixHybridProblem(y, rhs, lri.TwoFactorRepresentation(U0, Zh), (0.0, 0.04)), alg, 0.01)
        assert len(sol.Y) == 5
        assert max(rel(s.full(), g) for s, g in zip(sol.Y, G[f"hybrid_carry{int(carry)}_Y"])) <= 1e-10
        assert max(rel(s.Z, g) for s, g in zip(sol.Y, G[f"hybrid_carry{int(carry)}_Z"])) <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", DATA_ALGS)
def test_engine_matches_golden_data_problem(name):
    import lowrankintegrators.jl_b200 as lri
    algs = {"bug": lri.UnconventionalAlgorithm(), "ksl_primal": lri.ProjectorSplitting(lri.PrimalLieTrotter()),
            "ksl_dual": lri.ProjectorSplitting(lri.DualLieTrotter()), "rabug": lri.RankAdaptiveUnconventionalAlgorithm(1e-2, rmax=8),
            "greedy": lri.GreedyIntegrator()}
    u0 = lri.SVDLikeRepresentation(G["data_U0"], G["data_S0"], G["data_V0"])
    sol = lri.solve(lri.MatrixDataProblem(list(G["data_snaps"]), u0), algs[name])
    assert [y.rank for y in sol.Y] == list(G[f"data_{name}_rank"])
    assert max(rel(y.full(), g) for y, g in zip(sol.Y, G[f"data_{name}_Y"])) <= 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bug", "ksl_strang"])
def test_engine_matches_golden_de_problem(name):
    import torch
    import lowrankintegrators.jl_b200 as lri
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(np.asarray(x).T)).cuda().t()
    N = G["de_W1"].shape[0]
    rk4 = lri.SubStepper("rk4", nsub=2)
    algs = {"bug": lri.UnconventionalAlgorithm(K_alg=rk4, L_alg=rk4, S_alg=rk4),
            "ksl_strang": lri.ProjectorSplitting(lri.Strang(), K_alg=rk4, L_alg=rk4, S_alg=rk4)}
    rhs = lri.LinearRHS(A=dev(G["de_W1"] + np.eye(N)), B=dev(G["de_W2"].T))
    u0 = lri.SVDLikeRepresentation(G["de_U0"], G["de_S0"], G["de_V0"])
    sol = lri.solve(lri.MatrixDEProblem(rhs, u0, (0.0, 0.04)), algs[name], 0.01)
    assert len(sol.Y) == len(G[f"de_{name}_Y"])
    assert max(rel(y.full(), g) for y, g in zip(sol.Y, G[f"de_{name}_Y"])) <= 1e-10
