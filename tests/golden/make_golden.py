"""Generates tests/golden/oracle_golden.npz: seeded inputs and the ORACLE's per-step outputs for every integrator.
The reference itself holds no golden vectors and cannot run here (Julia absent), so these fixtures pin the
oracle restatement against drift and give the GPU parity tests a file-based target that needs no oracle import.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dlra_oracle as O  # noqa: E402
from tests.problems import lowrank_stream, skew_pair  # noqa: E402


def main():
    out = {}
    n, m, r, nsnap = 64, 48, 5, 4
    A = lowrank_stream(n, m, 9, seed=42, eps=1e-3)
    snaps = [A(0.07 * k) for k in range(nsnap)]
    X0 = O.truncated_svd(snaps[0], r)
    out["data_snaps"] = np.stack(snaps)
    out["data_U0"], out["data_S0"], out["data_V0"] = X0.U, X0.S, X0.V
    algs = {
        "bug": O.UnconventionalAlgorithm(),
        "ksl_primal": O.ProjectorSplitting(O.PrimalLieTrotter()),
        "ksl_dual": O.ProjectorSplitting(O.DualLieTrotter()),
        "rabug": O.RankAdaptiveUnconventionalAlgorithm(1e-2, rmax=8),
        "greedy": O.GreedyIntegrator(),
    }
    for name, alg in algs.items():
        sol = O.solve(O.MatrixDataProblem(snaps, X0), alg)
        out[f"data_{name}_Y"] = np.stack([y.full() for y in sol.Y])
        out[f"data_{name}_rank"] = np.array([y.rank for y in sol.Y])
    # DE problem: F(X) = W1 X + X + X W2 (linear), RK4 sub-steps
    N, rr = 48, 4
    W1, W2 = (0.2 * W for W in skew_pair(N, seed=43))
    D0 = np.diag(2.0 ** -np.arange(1, N + 1)) + 1e-3 * np.random.default_rng(44).standard_normal((N, N))
    Z0 = O.truncated_svd(D0, rr)
    out["de_W1"], out["de_W2"] = W1, W2
    out["de_U0"], out["de_S0"], out["de_V0"] = Z0.U, Z0.S, Z0.V
    f = lambda X, t: W1 @ X + X + X @ W2
    rk4 = O.SubStepper("rk4", nsub=2)
    for name, alg in {"bug": O.UnconventionalAlgorithm(K_alg=rk4, L_alg=rk4, S_alg=rk4),
                      "ksl_strang": O.ProjectorSplitting(O.Strang(), K_alg=rk4, L_alg=rk4, S_alg=rk4)}.items():
        sol = O.solve(O.MatrixDEProblem(f, Z0, (0.0, 0.04)), alg, 0.01)
        out[f"de_{name}_Y"] = np.stack([y.full() for y in sol.Y])
    # greedy steps on u = U*Z' (greedy_integrator.jl:72-92): data problem on the snapshots above, hybrid problem on the linear flow
    Zs = snaps[0].T @ X0.U
    sol = O.solve(O.MatrixDataProblem(snaps, O.TwoFactorRepresentation(X0.U, Zs)), O.GreedyIntegrator())
    out["data_greedy2_Y"] = np.stack([y.full() for y in sol.Y])
    from scipy.linalg import expm
    # the data rotate with the *other* generator, so that the basis update does not commute with the projected operator
    # U'AU of the Z-flow (otherwise the carried first stage and a fresh one coincide to round-off)
    Y0 = Z0.full()
    yfun = lambda t: expm(3.0 * t * W2) @ Y0 @ expm(3.0 * t * W1)
    fz = lambda Z, U, t: f(U @ Z.T, t).T @ U
    for carry in (True, False):
        alg = O.GreedyIntegrator(Z_alg=O.SubStepper("rk4", nsub=2), fsal_carry=carry)
        sol = O.solve(O.MatrixHybridProblem(yfun, fz, O.TwoFactorRepresentation(Z0.U, Y0.T @ Z0.U), (0.0, 0.04)), alg, 0.01)
        out[f"hybrid_carry{int(carry)}_Y"] = np.stack([y.full() for y in sol.Y])
        # for this linear F the carried stage only rotates Z (and U with it): U*Z' moves by 1e-10, Z itself by 1e-5
        out[f"hybrid_carry{int(carry)}_Z"] = np.stack([y.Z for y in sol.Y])
    out["hybrid_snaps"] = np.stack([yfun(0.01 * k) for k in range(5)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
