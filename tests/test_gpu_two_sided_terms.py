"""GPU parity of right-hand sides with two-sided terms F(X) = Σ_k A_k·X·B_kᵀ (dlra_rhs_add_term) against the CPU oracle:
random dense terms for every integrator and the chemical master equation of examples/markov_chain.jl at reduced size.

Validated on a B200 at the start of round 2 (16 cases green, gpurun_out/r2_base/pytest_optin.txt)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import dlra_oracle as O
from tests.problems import cme_operators, rel_fro
from tests.test_gpu_de_parity import algs, csr_dev, dev, run_both

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def lri():
    import torch
    assert torch.cuda.is_available()
    import lowrankintegrators.jl_b200 as lri
    return lri


@pytest.mark.parametrize("name", ["bug", "ksl_primal", "ksl_dual", "ksl_strang", "rabug"])
@pytest.mark.parametrize("sub", [("rk4", 2), ("tsit5", 1)])
def test_dense_two_sided_terms(lri, name, sub):
    n, m, r = 96, 80, 5
    rng = np.random.default_rng(11)
    As = [0.3 * rng.standard_normal((n, n)) / np.sqrt(n) for _ in range(3)]
    Bs = [0.3 * rng.standard_normal((m, m)) / np.sqrt(m) for _ in range(3)]
    A0 = 0.2 * rng.standard_normal((n, n)) / np.sqrt(n)
    X0 = O.truncated_svd((rng.standard_normal((n, r)) * 2.0 ** -np.arange(r)) @ rng.standard_normal((r, m)), r)
    of = lambda X, t: A0 @ X + sum(A @ X @ B.T for A, B in zip(As, Bs)) + 0.5 * X @ Bs[0].T + 0.1 * As[1] @ X
    # mixes the one-sided slots with two-sided terms, identity on either side included
    grhs = lri.FactoredRHS(A=dev(A0), terms=[(dev(A), dev(B)) for A, B in zip(As, Bs)] + [(0.5, dev(Bs[0])), (dev(As[1]), 0.1)])
    galg, oalg = algs(lri, lri.SubStepper(*sub), O.SubStepper(*sub))[name]
    errs = run_both(lri, grhs, of, X0, galg, oalg, 0.02, 3)
    assert max(errs) <= TOL, errs


@pytest.mark.parametrize("name", ["bug", "ksl_primal"])
def test_chemical_master_equation(lri, name):
    N, r = 64, 6
    terms = cme_operators(N)
    dense = [(A.toarray(), B.toarray()) for A, B in terms]
    of = lambda P, t: sum(A @ P @ B.T for A, B in dense)
    xs = np.arange(1, N + 1)
    D = np.array([[0.03, 0.01], [0.01, 0.02]])
    P0 = np.array([[np.exp(-np.array([x - 20, y - 20]) @ D @ np.array([x - 20, y - 20])) for y in xs] for x in xs])
    P0 /= P0.sum()
    X0 = O.truncated_svd(P0, r)
    grhs = lri.SylvesterSumRHS([(csr_dev(A), csr_dev(B)) for A, B in terms])
    sub = ("rk4", 2)
    galg, oalg = algs(lri, lri.SubStepper(*sub), O.SubStepper(*sub))[name]
    errs = run_both(lri, grhs, of, X0, galg, oalg, 2e-3, 4)
    assert max(errs) <= TOL, errs
