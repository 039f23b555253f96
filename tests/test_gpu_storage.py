"""Solution storage on the device path (primitives.jl:82-104): asynchronous `update_sol!` through
dlra_save_factors_async (pinned host buffers, copy stream), checkpoint -> resume reproducing the uninterrupted trajectory,
and the lifetime of borrowed device snapshots when the host runs many asynchronous steps ahead (engine.py)."""
import numpy as np
import pytest

from oracle import dlra_oracle as O
from tests.problems import lowrank_stream, rel_fro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lri():
    import torch
    assert torch.cuda.is_available()
    import lowrankintegrators.jl_b200 as lri
    return lri


def _dev(snaps):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(s.T)).cuda().t() for s in snaps]


def _algs(lri):
    return [lri.UnconventionalAlgorithm(), lri.ProjectorSplitting(lri.PrimalLieTrotter()),
            lri.RankAdaptiveUnconventionalAlgorithm(1e-6, rmax=16), lri.GreedyIntegrator()]


@pytest.mark.parametrize("ialg", range(4))
def test_async_save_equals_synchronous_save(lri, ialg):
    A = lowrank_stream(4096, 512, 12, seed=3, eps=1e-4)
    snaps = [A(0.04 * k) for k in range(9)]
    X0 = O.truncated_svd(snaps[0], 8)
    u0 = lri.SVDLikeRepresentation(X0.U, X0.S, X0.V)
    alg = _algs(lri)[ialg]
    ref = lri.solve(lri.MatrixDataProblem(_dev(snaps), u0), alg)
    got = lri.solve(lri.MatrixDataProblem(_dev(snaps), u0), alg, async_save=True)
    assert len(ref.Y) == len(got.Y) == len(snaps)
    assert ref.t == got.t
    for a, b in zip(ref.Y, got.Y):
        assert a.rank == b.rank
        for x, y in ((a.U, b.U), (a.S, b.S), (a.V, b.V)):
            assert np.array_equal(np.asarray(x), np.asarray(y))   # same device bytes, two download paths


@pytest.mark.parametrize("ialg", range(4))
def test_checkpoint_resume_reproduces_trajectory(lri, ialg):
    A = lowrank_stream(2048, 384, 12, seed=9, eps=1e-4)
    snaps = [A(0.04 * k) for k in range(9)]
    X0 = O.truncated_svd(snaps[0], 8)
    u0 = lri.SVDLikeRepresentation(X0.U, X0.S, X0.V)
    alg = _algs(lri)[ialg]
    full = lri.solve(lri.MatrixDataProblem(_dev(snaps), u0), alg)
    # interrupted run: 4 steps, checkpoint, a NEW engine resumes from the saved state and the stream position
    integ = lri.init(lri.MatrixDataProblem(_dev(snaps), u0), alg, 1)
    for _ in range(4):
        lri.step(integ)
    state = integ.checkpoint()
    integ.cache.close()
    assert state["t"] == 5 and state["iter"] == 4
    rest = lri.solve(lri.MatrixDataProblem(_dev(snaps), u0), alg, resume=state)
    assert len(rest.Y) == len(snaps) - 4
    for k, y in enumerate(rest.Y):
        ref = full.Y[4 + k]
        assert y.rank == ref.rank
        assert rel_fro(y.full(), ref.full()) <= 1e-12, (k, rel_fro(y.full(), ref.full()))


def test_fresh_snapshots_without_host_sync(lri):
    """y(t) returns a freshly allocated CUDA tensor every call and nothing synchronises the host (save_everystep=False):
    the host runs many steps ahead of the engine stream, the caller drops each snapshot right away and torch's caching
    allocator would hand the memory to the next y(t) while queued steps still read it — unless the engine registers its
    borrowed snapshots with its stream."""
    import torch
    n, m, r = 8192, 2048, 8
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    P = torch.rand((n, r), generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    Wm = torch.rand((m, r), generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    om = torch.linspace(0.5, 2.0, r, device="cuda", dtype=torch.float64)

    def y(t):   # exact rank r: BUG reproduces y(t_k) to round-off (exactness property) if it reads the right bytes
        out = lri.empty_colmajor(n, m, "cuda")
        out.copy_((P * torch.cos(om * t)) @ Wm.T)
        return out

    Y0 = y(0.0)
    Qp, Rp = torch.linalg.qr(P * torch.cos(om * 0.0))
    Qw, Rw = torch.linalg.qr(Wm)
    Us, s, Vh = torch.linalg.svd(Rp @ Rw.T)
    u0 = lri.SVDLikeRepresentation((Qp @ Us).cpu().numpy(), np.diag(s.cpu().numpy()), (Qw @ Vh.T).cpu().numpy())
    del Y0
    nsteps = 40
    for alg in (lri.UnconventionalAlgorithm(), lri.ProjectorSplitting(lri.PrimalLieTrotter())):
        sol = lri.solve(lri.MatrixDataProblem(y, u0, (0.0, 0.01 * nsteps)), alg, 0.01, save_everystep=False)
        final = sol.Y[-1]
        want = y(sol.t[-1]).cpu().numpy()
        assert rel_fro(final.full(), want) < 1e-11
