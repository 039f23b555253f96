"""GPU parity at the factor widths the BASELINE configs need: r = 48 / 64 for BUG and KSL (cfg 5 —
thread-block clusters of 3 / 4 CTAs sharing each ΔA tile through TMA multicast, 16 factor columns per CTA) and the
rank-adaptive step with augmented widths 2r in {128, 160, 256} (cfg 4: the L2-resident global-memory path of the
one-block Jacobi SVD, csrc/jacobi.cuh, and 128-column K-only sweeps).  Reference lines: data_integrator.jl:13-16,
unconventional.jl:133-157, projector_splitting.jl:129-189, rank_adaptive_unconventional.jl:194-233."""
import numpy as np
import pytest

from oracle import dlra_oracle as O
from tests.problems import lowrank_stream, rel_fro
from tests.test_gpu_data_parity import TOL, _run_both

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lri():
    import torch
    assert torch.cuda.is_available()
    import lowrankintegrators.jl_b200 as lri
    return lri


@pytest.mark.parametrize("name", ["bug", "ksl_primal", "ksl_dual", "greedy"])
@pytest.mark.parametrize("shape", [(4096, 512, 64), (2048, 640, 48), (8192, 256, 64), (3072, 384, 32), (16384, 320, 64), (12288, 256, 40)])
def test_step_parity_wide(lri, name, shape):
    n, m, r = shape
    A = lowrank_stream(n, m, 2 * r, seed=17, eps=1e-4)
    snaps = [A(0.03 * k) for k in range(4)]
    errs = _run_both(lri, name, snaps, r, True, resync=True)
    assert max(errs) <= TOL, errs
    errs = _run_both(lri, name, snaps, r, True)
    assert max(errs) <= TOL, errs


@pytest.mark.parametrize("r0", [64, 80, 128])
def test_rabug_wide_core(lri, r0):
    # augmented core of 2*r0 in {128, 160, 256} columns: Jacobi SVD from the global-memory scratch, BCGS2 over up to 16 panels
    import torch
    n, m = 4096, 512
    R = r0 + r0 // 4
    rng = np.random.default_rng(3)
    Pn, Wn = rng.uniform(-1, 1, (n, R)), rng.uniform(-1, 1, (m, R))
    sig = 1.2 ** -np.arange(R)   # slower decay than lowrank_stream's 2^-q: every one of the R directions stays above round-off

    def Y(t):
        return (Pn * (sig * np.cos(0.7 * t + np.arange(R)))) @ Wn.T

    snaps = [Y(0.05 * k) for k in range(3)]
    X0 = O.truncated_svd(snaps[0], r0)
    galg = lri.RankAdaptiveUnconventionalAlgorithm(1e-7, rmax=128)
    oalg = O.RankAdaptiveUnconventionalAlgorithm(1e-7, rmax=128)
    oint = O.init(O.MatrixDataProblem(snaps, X0), oalg, 1)
    dsnaps = [torch.from_numpy(np.ascontiguousarray(s.T)).cuda().t() for s in snaps]
    gint = lri.init(lri.MatrixDataProblem(dsnaps, lri.SVDLikeRepresentation(X0.U, X0.S, X0.V)), galg, 1)
    for k in range(len(snaps) - 1):
        gint.cache.set_factors(oint.u.U, oint.u.S, oint.u.V)   # per-step parity from identical inputs
        O.step(oint)
        lri.step(gint)
        gu, ou = gint.u, oint.u
        assert gu.rank == ou.rank, (k, gu.rank, ou.rank)
        assert rel_fro(gu.full(), ou.full()) <= TOL
        assert np.linalg.norm(gu.U.T @ gu.U - np.eye(gu.rank)) < 1e-12
        assert np.linalg.norm(gu.V.T @ gu.V - np.eye(gu.rank)) < 1e-12
