"""GPU parity of the data-problem steps against the CPU oracle, through the C ABI (ctypes -> libdlra.so).
Bar (BASELINE.json north_star): relative Frobenius error of the reconstructed U·S·Vᵀ <= 1e-10 per step,
identical selected ranks for the rank-adaptive path."""
import numpy as np
import pytest

from oracle import dlra_oracle as O
from tests.problems import generic_matrix_stream, lowrank_stream, rel_fro

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def lri():
    import torch
    assert torch.cuda.is_available()
    import lowrankintegrators.jl_b200 as lri
    return lri


def _algs(lri):
    return {
        "bug": (lri.UnconventionalAlgorithm(), O.UnconventionalAlgorithm()),
        "ksl_primal": (lri.ProjectorSplitting(lri.PrimalLieTrotter()), O.ProjectorSplitting(O.PrimalLieTrotter())),
        "ksl_dual": (lri.ProjectorSplitting(lri.DualLieTrotter()), O.ProjectorSplitting(O.DualLieTrotter())),
        "rabug": (lri.RankAdaptiveUnconventionalAlgorithm(1e-6, rmax=12), O.RankAdaptiveUnconventionalAlgorithm(1e-6, rmax=12)),
        "greedy": (lri.GreedyIntegrator(), O.GreedyIntegrator()),
    }


def _run_both(lri, name, snaps, r0, device_data, force_generic=False, nsteps=None, resync=False):
    import torch
    X0 = O.truncated_svd(snaps[0], r0)
    galg, oalg = _algs(lri)[name]
    oint = O.init(O.MatrixDataProblem(snaps, X0), oalg, 1)
    if device_data:
        dsnaps = [torch.from_numpy(np.ascontiguousarray(s.T)).cuda().t() for s in snaps]  # column-major on device
    else:
        dsnaps = snaps
    gu0 = lri.SVDLikeRepresentation(X0.U, X0.S, X0.V)
    gint = lri.init(lri.MatrixDataProblem(dsnaps, gu0), galg, 1, force_generic=force_generic)
    errs = []
    for k in range(nsteps or len(snaps) - 1):
        if resync:  # per-step parity from IDENTICAL inputs (the north star's "per step" bar)
            gint.cache.set_factors(oint.u.U, oint.u.S, oint.u.V)
        O.step(oint)
        lri.step(gint)
        gu, ou = gint.u, oint.u
        assert gu.rank == ou.rank, f"rank mismatch at step {k}: {gu.rank} vs {ou.rank}"
        errs.append(rel_fro(gu.full(), ou.full()))
        # factors must be orthonormal like the reference's
        assert np.linalg.norm(gu.U.T @ gu.U - np.eye(gu.rank)) < 1e-12
        assert np.linalg.norm(gu.V.T @ gu.V - np.eye(gu.rank)) < 1e-12
    return errs


@pytest.mark.parametrize("name", ["bug", "ksl_primal", "ksl_dual", "rabug", "greedy"])
@pytest.mark.parametrize("device_data", [False, True])
def test_step_parity_small_generic(lri, name, device_data):
    # odd sizes: generic kernels.  The adaptive case truncates inside a clear spectral gap (exact rank 10): cutting
    # through a cluster of noise singular values is ill-conditioned for ANY implementation (the oracle itself moves
    # by 1e-10 under a 4e-16 relative input perturbation there).
    A = lowrank_stream(301, 203, 10, seed=5, eps=0.0 if name == "rabug" else 1e-3)
    snaps = [A(0.05 * k) for k in range(6)]
    errs = _run_both(lri, name, snaps, 6, device_data, force_generic=True)
    assert max(errs) <= TOL, errs
    errs = _run_both(lri, name, snaps, 6, device_data, force_generic=True, resync=True)
    assert max(errs) <= TOL, errs


@pytest.mark.parametrize("name", ["bug", "ksl_primal", "ksl_dual", "rabug", "greedy"])
@pytest.mark.parametrize("shape", [(1024, 512, 16), (2048, 384, 8), (4096, 256, 32), (1000, 130, 5), (2048, 256, 24), (1536, 320, 40)])
def test_step_parity_fast_path(lri, name, shape):
    n, m, r = shape
    A = lowrank_stream(n, m, 2 * r if name != "rabug" else r + r // 2, seed=7, eps=0.0 if name == "rabug" else 1e-4)
    snaps = [A(0.03 * k) for k in range(4)]
    errs = _run_both(lri, name, snaps, r, True, resync=True)
    assert max(errs) <= TOL, errs
    errs = _run_both(lri, name, snaps, r, True)
    assert max(errs) <= TOL, errs


def test_reference_selfconsistency_on_gpu(lri):
    # test/data_driven_approximation.jl:2-30 through the GPU engine: continuous stream == discrete snapshots
    Y = generic_matrix_stream(100, seed=0)
    X0 = O.truncated_svd(Y(0.0), tol=1e-4)
    u0 = lri.SVDLikeRepresentation(X0.U, X0.S, X0.V)
    data = [Y(t) for t in np.arange(0, 1.0 + 1e-12, 0.01)]
    for galg, oalg in _algs(lri).values():
        if isinstance(galg, lri.RankAdaptiveUnconventionalAlgorithm):
            galg, oalg = lri.RankAdaptiveUnconventionalAlgorithm(1e-8, rmax=20), O.RankAdaptiveUnconventionalAlgorithm(1e-8, rmax=20)
        sol = lri.solve(lri.MatrixDataProblem(Y, u0, (0.0, 1.0)), galg, 1e-2)
        dsol = lri.solve(lri.MatrixDataProblem(data, u0), galg)
        assert len(dsol.Y) == 101
        assert np.allclose(sol.Y[-1].full(), dsol.Y[-1].full(), rtol=np.sqrt(np.finfo(float).eps), atol=1e-12)
        osol = O.solve(O.MatrixDataProblem(data, X0), oalg)
        assert [y.rank for y in dsol.Y] == [y.rank for y in osol.Y]
        assert rel_fro(dsol.Y[-1].full(), osol.Y[-1].full()) <= 1e-9  # 100 accumulated steps


def test_exactness_on_rank_r_stream_gpu(lri):
    A = lowrank_stream(2048, 512, 8, seed=11)
    snaps = [A(0.05 * k) for k in range(11)]
    X0 = O.truncated_svd(snaps[0], 8)
    u0 = lri.SVDLikeRepresentation(X0.U, X0.S, X0.V)
    for alg in (lri.UnconventionalAlgorithm(), lri.ProjectorSplitting(lri.PrimalLieTrotter()), lri.ProjectorSplitting(lri.DualLieTrotter())):
        sol = lri.solve(lri.MatrixDataProblem(snaps, u0), alg, save_everystep=False)
        assert rel_fro(sol.Y[-1].full(), snaps[-1]) < 1e-12


def test_strang_on_snapshot_vector_is_method_error(lri):
    A = lowrank_stream(64, 32, 3, seed=2)
    snaps = [A(0.1 * k) for k in range(4)]
    X0 = O.truncated_svd(snaps[0], 3)
    u0 = lri.SVDLikeRepresentation(X0.U, X0.S, X0.V)
    with pytest.raises(TypeError):
        lri.solve(lri.MatrixDataProblem(snaps, u0), lri.ProjectorSplitting(lri.Strang()))
    Y = lowrank_stream(64, 32, 3, seed=2)
    sol = lri.solve(lri.MatrixDataProblem(Y, u0, (0.0, 0.3)), lri.ProjectorSplitting(lri.Strang()), 0.1)
    osol = O.solve(O.MatrixDataProblem(Y, X0, (0.0, 0.3)), O.ProjectorSplitting(O.Strang()), 0.1)
    assert rel_fro(sol.Y[-1].full(), osol.Y[-1].full()) <= TOL


def test_errors_are_loud(lri):
    eng = lri.Engine(64, 32, 4)
    with pytest.raises(lri._lib.DLRAError):
        eng.step_bug()          # no data pushed
    with pytest.raises(lri._lib.DLRAError):
        lri.Engine(64, 32, 4, rmax=2)


def test_truncated_svd_device(lri):
    # initial condition without a host SVD of n x m.  Exact for rank(A) <= r, near-optimal otherwise.
    import torch
    rng = np.random.default_rng(5)
    n, m = 3000, 700
    Q1, _ = np.linalg.qr(rng.standard_normal((n, 40)))
    Q2, _ = np.linalg.qr(rng.standard_normal((m, 40)))
    sig = 2.0 ** -np.arange(40)
    A = (Q1 * sig) @ Q2.T
    Ad = torch.from_numpy(np.ascontiguousarray(A.T)).cuda().t()
    for r in (6, 16, 24):
        u = lri.truncated_svd_device(Ad, r, oversample=8, power_iters=2)
        best = np.sqrt(np.sum(sig[r:] ** 2))
        err = np.linalg.norm(u.full() - A)
        assert u.rank == r and err <= 1.02 * best, (r, err, best)
        assert np.linalg.norm(u.U.T @ u.U - np.eye(r)) < 1e-12 and np.linalg.norm(u.V.T @ u.V - np.eye(r)) < 1e-12
        assert np.allclose(np.diag(u.S), sig[:r], rtol=1e-3)
    u = lri.truncated_svd_device(Ad, tol=1e-4, rmax=32)
    ref = O.truncated_svd(A, tol=1e-4)
    assert u.rank == ref.rank
    B = (Q1[:, :5] * sig[:5]) @ Q2[:, :5].T           # exactly rank 5: reproduced to round-off
    Bd = torch.from_numpy(np.ascontiguousarray(B.T)).cuda().t()
    u = lri.truncated_svd_device(Bd, 5)
    assert rel_fro(u.full(), B) < 1e-13


def test_switching_integrators_after_pipelined_bug_steps(lri):
    # The pipelined BUG step leaves work for the NEXT step queued on the auxiliary stream (sum of the L partials into the spare V buffer,
    # M = U1'U0 reading the old basis).  A different integrator that follows must see none of it: the KSL steps
    # reuse those buffers on the main stream.  Shapes on the TMA fast path (lookahead on by default).
    import torch
    n, m, r = 4096, 512, 12
    A = lowrank_stream(n, m, r, seed=11, eps=0.0)
    snaps = [A(0.03 * k) for k in range(9)]
    dsnaps = [torch.from_numpy(np.ascontiguousarray(s.T)).cuda().t() for s in snaps]
    X0 = O.truncated_svd(snaps[0], r)
    # (no greedy step in the sequence: the reference's greedy step does not advance `yprev`, so what follows it is not comparable)
    galgs, oalgs = zip(*[_algs(lri)[k] for k in ("bug", "bug", "ksl_primal", "bug", "bug", "ksl_dual", "bug", "ksl_primal")])
    oint = O.init(O.MatrixDataProblem(snaps, X0), oalgs[0], 1)
    gint = lri.init(lri.MatrixDataProblem(dsnaps, lri.SVDLikeRepresentation(X0.U, X0.S, X0.V)), galgs[0], 1)
    for k, (galg, oalg) in enumerate(zip(galgs, oalgs)):
        O.step(oint, oalg)
        lri.step(gint, galg)
        gu, ou = gint.u, oint.u
        assert rel_fro(gu.full(), ou.full()) <= TOL, (k, type(galg).__name__)
        assert np.linalg.norm(gu.U.T @ gu.U - np.eye(gu.rank)) < 1e-12
        assert np.linalg.norm(gu.V.T @ gu.V - np.eye(gu.rank)) < 1e-12
