"""Device-evaluable right-hand sides for MatrixDEProblem (north star: linear operators A·X + X·Bᵀ, low-rank forcing,
and column-wise elementwise nonlinearities as in the Burgers UQ example).  They replace the Julia closure `prob.f`
evaluated through LowRankArithmetic (projector_splitting.jl:52-80): the engine evaluates the projected K/L/S
right-hand sides on the factors directly and never forms an n x m matrix."""
from dataclasses import dataclass
from typing import Optional


@dataclass
class FactoredRHS:
    """F(X, t) = A·X + X·Bᵀ + G·Hᵀ + c_had·(D1·X) .* (D2·X) + Σ_k A_k·X·B_kᵀ; every term optional.
    A, D1, D2: n x n, B: m x m — each a CUDA fp64 tensor (dense), a CSR tuple (rowptr, colind, values, shape) of CUDA
    tensors, or a Python scalar s meaning s·I.  G: n x q, H: m x q CUDA tensors."""
    A: object = None
    B: object = None
    G: object = None
    H: object = None
    D1: object = None
    D2: object = None
    c_had: float = 0.0
    terms: Optional[list] = None   # [(A_k, B_k), ...] two-sided terms A_k·X·B_kᵀ (operators as above; 1.0 for the identity)

    def install(self, engine):
        engine.rhs_set(self.A, self.B, self.G, self.H, self.D1, self.D2, self.c_had)
        for Ak, Bk in (self.terms or []):
            engine.rhs_add_term(Ak, Bk)


def LinearRHS(A=None, B=None, G=None, H=None):
    """F(X) = A·X + X·Bᵀ (+ G·Hᵀ): examples/generic_matrix.jl-style and Lyapunov-type problems."""
    return FactoredRHS(A=A, B=B, G=G, H=H)


def SylvesterSumRHS(terms):
    """F(X) = Σ_k A_k·X·B_kᵀ: e.g. the chemical master equation of examples/markov_chain.jl:64-66,
    Σ_r A_r .* (S_r·P·T_r) − Asum .* P with rank-one weights A_r = a_r·b_rᵀ, i.e. A_k = diag(a_r)·S_r, B_kᵀ = T_r·diag(b_r)."""
    return FactoredRHS(terms=list(terms))


def BurgersRHS(lap, grad):
    """F(ρ) = Δρ − (∇ρ) .* ρ  (test/data_agnostic_approximation.jl:31-33); lap/grad as accepted by FactoredRHS."""
    return FactoredRHS(A=lap, D1=grad, D2=1.0, c_had=-1.0)
