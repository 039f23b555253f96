"""Rank-adaptive step with the augmented bases factored as [U0 | K], [V0 | L] (DLRA_AUG_BASIS_FIRST): same span as the
reference's [K | U0], [L | V0], hence the same U·S·Vᵀ and the same selected ranks, with one TSQR less per side.

Validated on a B200 at the start of round 2 (16 cases green, gpurun_out/r2_base/pytest_optin.txt)."""
import numpy as np
import pytest

from oracle import dlra_oracle as O
from tests.problems import lowrank_stream, rel_fro

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,r,R", [(4096, 512, 16, 20), (2048, 384, 6, 10), (4096, 256, 24, 30), (2050, 130, 5, 8)])
def test_rabug_data_matches_oracle_with_basis_first(n, m, r, R):
    import lowrankintegrators.jl_b200 as lri
    A = lowrank_stream(n, m, R, seed=31, eps=0.0)      # exact rank R: the truncation does not cut through a noise cluster
    snaps = [A(0.04 * k) for k in range(5)]
    X0 = O.truncated_svd(snaps[0], r)
    galg = lri.RankAdaptiveUnconventionalAlgorithm(1e-6, rmax=2 * r)
    oalg = O.RankAdaptiveUnconventionalAlgorithm(1e-6, rmax=2 * r)
    gint = lri.init(lri.MatrixDataProblem(snaps, lri.SVDLikeRepresentation(X0.U, X0.S, X0.V)), galg, 1, aug_basis_first=True)
    oint = O.init(O.MatrixDataProblem(snaps, X0), oalg, 1)
    for k in range(4):
        O.step(oint)
        lri.step(gint)
        gu, ou = gint.u, oint.u
        assert gu.rank == ou.rank, (k, gu.rank, ou.rank)
        assert rel_fro(gu.full(), ou.full()) <= 1e-10, k
        assert np.linalg.norm(gu.U.T @ gu.U - np.eye(gu.rank)) < 1e-12
