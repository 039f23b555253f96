"""Seeded synthetic inputs shared by the oracle tests and the GPU parity tests."""
import numpy as np
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
