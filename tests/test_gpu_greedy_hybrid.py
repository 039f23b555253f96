"""GPU parity of the greedy integrator on u = U*Z' (TwoFactorRepresentation data problems and MatrixHybridProblem,
greedy_integrator.jl:72-92) and of normal_component (utils.jl:2-20) against the CPU oracle, through the C ABI
(SURVEY.md 8f item 4).  Bar: rel. Frobenius error of U*Z' <= 1e-10 per step."""
import numpy as np
import pytest

from oracle import dlra_oracle as O
from tests.problems import burgers_truth, lowrank_stream, rel_fro
from tests.test_gpu_de_parity import csr_dev, dev, periodic_ops

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def lri():
    import torch
    assert torch.cuda.is_available()
    import lowrankintegrators.jl_b200 as lri
    return lri


def two_factor_start(A0, r):
    U0 = np.linalg.svd(A0, full_matrices=False)[0][:, :r]
    return U0, A0.T @ U0


@pytest.mark.parametrize("n,m,r", [(384, 256, 6), (4096, 512, 8), (2050, 130, 5)])
def test_two_factor_greedy_data(lri, n, m, r):
    # generic kernels (first and last shape) and the TMA/DMMA pass kernels (middle shape)
    A = lowrank_stream(n, m, 2 * r, seed=2, eps=1e-6)
    snaps = [A(0.05 * k) for k in range(5)]
    U0, Z0 = two_factor_start(snaps[0], r)
    gint = lri.init(lri.MatrixDataProblem(snaps, lri.TwoFactorRepresentation(U0, Z0)), lri.GreedyIntegrator(), 1)
    oint = O.init(O.MatrixDataProblem(snaps, O.TwoFactorRepresentation(U0, Z0)), O.GreedyIntegrator(), 1)
    for k in range(4):
        O.step(oint)
        lri.step(gint)
        gu, ou = gint.u, oint.u
        assert isinstance(gu, lri.TwoFactorRepresentation)
        assert rel_fro(gu.full(), ou.full()) <= TOL, k
        assert rel_fro(gu.Z, ou.Z) <= TOL and np.linalg.norm(gu.U.T @ gu.U - np.eye(r)) < 1e-12


@pytest.mark.parametrize("n,m,r,deficit", [(384, 256, 6, 2), (4096, 512, 8, 3)])
def test_two_factor_greedy_rank_deficient_data(lri, n, m, r, deficit):
    # rank(X) < r: svd(X*Z) has zero singular values; the reference's polar factor Q*P' is orthonormal all the same
    # (U*Z' does not depend on the completion).  The one-sided Jacobi core SVD leaves the null-space columns of P non-orthogonal -> ortho_complete.
    rng = np.random.default_rng(5)
    s = r - deficit
    # The column space of the stream is FIXED: then X'q = 0 for every completion vector q of the polar factor, and Z, X*Z and U*Z'
    # of all later steps are independent of the (arbitrary) completion -- with a moving column space parity is only defined for step 1.
    L0, R0, R1 = rng.standard_normal((n, s)), rng.standard_normal((m, s)), rng.standard_normal((m, s))
    snaps = [L0 @ (R0 + 0.05 * k * R1).T for k in range(5)]
    U0, Z0 = two_factor_start(snaps[0], r)
    gint = lri.init(lri.MatrixDataProblem(snaps, lri.TwoFactorRepresentation(U0, Z0)), lri.GreedyIntegrator(), 1)
    oint = O.init(O.MatrixDataProblem(snaps, O.TwoFactorRepresentation(U0, Z0)), O.GreedyIntegrator(), 1)
    for k in range(4):
        O.step(oint)
        lri.step(gint)
        gu, ou = gint.u, oint.u
        assert np.linalg.norm(gu.U.T @ gu.U - np.eye(r)) < 1e-12, k
        assert rel_fro(gu.full(), ou.full()) <= TOL, (k, rel_fro(gu.full(), ou.full()))
        assert rel_fro(gu.full(), snaps[k + 1]) <= TOL, k     # span(U) contains the column space: U*U'*X = X


def hybrid_pair(lri, y, grhs, of, U0, Z0, tf, dt, sub, carry):
    galg = lri.GreedyIntegrator(Z_alg=lri.SubStepper(*sub), fsal_carry=carry)
    oalg = O.GreedyIntegrator(Z_alg=O.SubStepper(*sub), fsal_carry=carry)
    gint = lri.init(lri.MatrixHybridProblem(y, grhs, lri.TwoFactorRepresentation(U0, Z0), (0.0, tf)), galg, dt)
    oint = O.init(O.MatrixHybridProblem(y, lambda Z, U, t: of(U @ Z.T, t).T @ U, O.TwoFactorRepresentation(U0, Z0), (0.0, tf)), oalg, dt)
    return gint, oint


@pytest.mark.parametrize("sub", [("rk4", 2), ("euler", 3), ("tsit5_fixed", 1), ("tsit5", 1)])
@pytest.mark.parametrize("carry", [True, False])
def test_hybrid_linear_flow(lri, sub, carry):
    from scipy.linalg import expm
    n, m, r = 96, 80, 5
    rng = np.random.default_rng(1)
    A = rng.standard_normal((n, n)); A = 0.2 * (A - A.T)
    B = rng.standard_normal((m, m)); B = 0.2 * (B - B.T)
    Y0 = (rng.standard_normal((n, r)) * 2.0 ** -np.arange(r)) @ rng.standard_normal((r, m))
    Y0 = Y0 + 1e-4 * rng.standard_normal((n, m))
    y = lambda t: expm(t * A) @ Y0 @ expm(t * B.T)
    of = lambda X, t: A @ X + X @ B.T
    U0, Z0 = two_factor_start(Y0, r)
    gint, oint = hybrid_pair(lri, y, lri.LinearRHS(A=dev(A), B=dev(B)), of, U0, Z0, 0.2, 0.05, sub, carry)
    for k in range(4):
        O.step(oint)
        lri.step(gint)
        assert rel_fro(gint.u.full(), oint.u.full()) <= TOL, (k, sub, carry)


@pytest.mark.parametrize("carry", [True, False])
def test_hybrid_burgers_data_informed(lri, carry):
    # test/data_informed_approximation.jl at reduced size: CSR operators, column-wise nonlinearity, adaptive Tsit5 for Z
    n, mm, r, dt = 128, 6, 10, 1e-2
    xi = [(a, b) for b in np.linspace(-1, 1, mm) for a in np.linspace(-1, 1, mm)]
    t_grid = np.arange(0, 0.05 + 1e-12, dt)
    truth, _ = burgers_truth(n, xi, t_grid, nu=0.02)
    lap, grad = periodic_ops(n, nu=0.02)
    Ld, Gd = lap.toarray(), grad.toarray()
    of = lambda X, t: Ld @ X - (Gd @ X) * X
    U0 = O.truncated_svd(np.hstack([truth[0], truth[1]]), r).U
    Z0 = truth[0].T @ U0
    y = lambda t: truth[min(int(np.floor(t / dt + 1e-9)), len(t_grid) - 1)]
    gint, oint = hybrid_pair(lri, y, lri.BurgersRHS(csr_dev(lap), csr_dev(grad)), of, U0, Z0, 0.05, dt, ("tsit5", 1), carry)
    for k in range(5):
        O.step(oint)
        lri.step(gint)
        assert rel_fro(gint.u.full(), oint.u.full()) <= TOL, (k, carry)
    assert rel_fro(gint.u.full(), truth[-1]) <= 0.1


@pytest.mark.parametrize("n,m,r", [(300, 200, 4), (4096, 512, 16)])
@pytest.mark.parametrize("given_C", [False, True])
def test_normal_component(lri, n, m, r, given_C):
    rng = np.random.default_rng(7)
    U = np.linalg.qr(rng.standard_normal((n, r)))[0]
    Z = rng.standard_normal((m, r)) * 2.0 ** -np.arange(r)
    dY = rng.standard_normal((n, m))
    Cm = Z.T @ Z + 0.1 * np.eye(r) if given_C else None
    ref = O.normal_component(U, Z, dY, C=Cm)
    eng = lri.Engine(n, m, r)
    eng.set_factors(U, np.eye(r), Z)
    nrm, N = lri.normal_component(eng, dev(dY), C=None if Cm is None else dev(Cm), want_matrix=True)
    assert abs(nrm - np.linalg.norm(ref)) <= 1e-12 * np.linalg.norm(ref)
    assert rel_fro(N.cpu().numpy(), ref) <= 1e-12
    # SVD-like factors: Z = V*S'
    Q, R = np.linalg.qr(Z)
    eng.set_factors(U, R.T.copy(), Q)
    assert abs(eng.normal_component(dev(dY), None if Cm is None else dev(Cm)) - np.linalg.norm(ref)) <= 1e-11 * np.linalg.norm(ref)
    # rank-deficient Z: pinv(C, atol = tol) drops the null direction
    Z2 = Z.copy(); Z2[:, -1] = Z2[:, -2]
    eng.set_factors(U, np.eye(r), Z2)
    ref2 = O.normal_component(U, Z2, dY)
    assert abs(eng.normal_component(dev(dY)) - np.linalg.norm(ref2)) <= 1e-10 * np.linalg.norm(ref2)
    eng.close()
