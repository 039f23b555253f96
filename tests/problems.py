"""Seeded synthetic inputs shared by the oracle tests and the GPU parity tests."""
import numpy as np
import scipy.sparse as sp
from scipy.linalg import expm


def skew_pair(N, seed=0):
    """examples/generic_matrix.jl:2-13: two skew-symmetric matrices from N(0,1) upper triangles."""
    rng = np.random.default_rng(seed)
    Ws = []
    for _ in range(2):
        W = np.triu(rng.standard_normal((N, N)), 1)
        Ws.append(W - W.T)
    return Ws


def generic_matrix_stream(N=100, seed=0):
    """Y(t) = exp(t W1) * e^t D * exp(t W2), D = diag(2^-j) (examples/generic_matrix.jl:15-18)."""
    W1, W2 = skew_pair(N, seed)
    D = np.diag(2.0 ** -np.arange(1, N + 1))
    return lambda t: expm(t * W1) @ (np.exp(t) * D) @ expm(t * W2)


def lowrank_stream(n, m, R, seed=0, eps=0.0):
    """A(t) = P diag(sigma_q cos(w_q t + phi_q)) W' + eps*H(t)  (SURVEY.md 8d config 2 recipe)."""
    rng = np.random.default_rng(seed)
    P = rng.uniform(-1, 1, (n, R))
    W = rng.uniform(-1, 1, (m, R))
    sig = 2.0 ** -np.arange(R)
    om = rng.uniform(0.5, 2.0, R)
    ph = rng.uniform(0, 2 * np.pi, R)
    H = rng.uniform(-1, 1, (n, m)) if eps else None

    def A(t):
        out = (P * (sig * np.cos(om * t + ph))) @ W.T
        if eps:
            out = out + eps * np.cos(3.0 * t) * H
        return out
    return A


def rel_fro(X, Y):
    return np.linalg.norm(X - Y) / np.linalg.norm(Y)


def burgers_truth(n, xi_grid, t_grid, nu=0.005, length=np.pi, substeps=20):
    """Column-wise Burgers solves standing in for the 100 ODE solves of test/data_informed_approximation.jl:41-64
    (classical RK4 with `substeps` sub-steps per output interval instead of DifferentialEquations.solve)."""
    dx = length / n
    x = (np.arange(n) + 0.5) * dx
    ub = 0.5 * (np.exp(np.cos(x)) - 1.5) * np.sin(x + 2 * np.pi * 0.37)
    rho = np.stack([ub + 0.5 * a * np.sin(2 * np.pi * x) + 0.5 * b * np.sin(3 * np.pi * x) for a, b in xi_grid], axis=1)

    def F(r):
        lap = (np.roll(r, 1, 0) - 2 * r + np.roll(r, -1, 0)) * (nu / dx ** 2)
        grad = (np.roll(r, -1, 0) - np.roll(r, 1, 0)) * (0.5 / dx)
        return lap - grad * r

    out = [rho.copy()]
    for k in range(1, len(t_grid)):
        h = (t_grid[k] - t_grid[k - 1]) / substeps
        for _ in range(substeps):
            k1 = F(rho); k2 = F(rho + 0.5 * h * k1); k3 = F(rho + 0.5 * h * k2); k4 = F(rho + h * k3)
            rho = rho + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        out.append(rho.copy())
    return out, F


def cme_operators(N, k1=30.0, k2=1.0, k3=10.0, k4=1.0, theta=1.0):
    """examples/markov_chain.jl:10-66: propensity factors ax, ay and shift operators of the two-species network."""
    x = np.arange(N, dtype=np.float64)
    ax = [np.ones(N), k2 * x, k3 / (1 + x), np.ones(N)]
    ay = [k1 / (1 + (x / theta) ** 3), np.ones(N), np.ones(N), k4 * x]
    nu = [(1, 0), (-1, 0), (0, 1), (0, -1)]

    def shift(s):
        return sp.eye(N, k=-s, format="csr") if s != 0 else sp.identity(N, format="csr")
    terms = []
    for r in range(4):
        Sr, Sc = shift(nu[r][0]), shift(nu[r][1])              # Srows[r], Scols[r]' = shift(nu_y)'
        a_sh, b_sh = Sr @ ax[r], Sc @ ay[r]                      # A[r] = Srows*TwoFactor(ax,ay)*Scols: shifted rank-one weights
        # A[r] .* (Srows P Scols) = diag(a_sh)·Srows·P·Scols·diag(b_sh)
        terms.append((sp.csr_matrix(sp.diags(a_sh) @ Sr), sp.csr_matrix(sp.diags(b_sh) @ Sc)))   # (A_k, B_k) with B_kᵀ = Scols·diag
        terms.append((sp.csr_matrix(-sp.diags(ax[r])), sp.csr_matrix(sp.diags(ay[r]))))           # − Asum .* P
    return terms
