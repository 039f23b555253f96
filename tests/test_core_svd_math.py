"""CPU pins for the arithmetic the core-SVD kernel (csrc/jacobi.cuh) relies on, restated in NumPy:
 * the rotation written with two rsqrt and one division equals the textbook zeta/t/c/s form,
 * one-sided Jacobi on the transposed triangular factor R' of A = Q0*R, with the accumulator started at Q0, yields
   A = V * diag(sigma) * (G/sigma)' (accumulator = LEFT vectors, normalised columns = RIGHT vectors), and needs far fewer sweeps than
   Jacobi on A itself when the singular values are graded (the claim behind the default for 2r >= 64)."""
import numpy as np


def _sweeps(G0, V0, eps_rot, new_rotation=True, max_sweeps=60):
    N = G0.shape[0]
    G, V = G0.copy(), V0.copy()
    noise2 = 1e-30 * np.sum(G0 * G0)
    Np = (N + 1) & ~1
    for sweep in range(max_sweeps):
        rotated = 0
        for rnd in range(Np - 1):
            for pi in range(Np // 2):   # round-robin schedule of jacobi_cluster_kernel (super-round 0)
                p, q = (Np - 1, rnd) if pi == 0 else ((rnd + pi) % (Np - 1), (rnd - pi + (Np - 1)) % (Np - 1))
                if p >= N or q >= N:
                    continue
                if p > q:
                    p, q = q, p
                a, b, g = G[:, p] @ G[:, p], G[:, q] @ G[:, q], G[:, p] @ G[:, q]
                ab = a * b
                if not (ab > noise2 * noise2 and g * g > eps_rot * eps_rot * ab):
                    continue
                if new_rotation:
                    d = b - a
                    x = d * d + 4.0 * g * g
                    t = (2.0 * g if d >= 0.0 else -2.0 * g) / (abs(d) + x * (1.0 / np.sqrt(x)))
                    c = 1.0 / np.sqrt(t * t + 1.0)
                else:
                    zeta = (b - a) / (2.0 * g)
                    t = np.copysign(1.0, zeta) / (abs(zeta) + np.sqrt(1.0 + zeta * zeta))
                    c = 1.0 / np.sqrt(1.0 + t * t)
                s = c * t
                rotated += 1
                for M in (G, V):
                    x_, y_ = M[:, p].copy(), M[:, q].copy()
                    M[:, p], M[:, q] = c * x_ - s * y_, s * x_ + c * y_
        if rotated == 0:
            return sweep + 1, G, V
    return max_sweeps, G, V


def _graded(N, base, rng):
    Q1 = np.linalg.qr(rng.standard_normal((N, N)))[0]
    Q2 = np.linalg.qr(rng.standard_normal((N, N)))[0]
    return Q1 @ np.diag(base ** -np.arange(N, dtype=float)) @ Q2.T


def test_rotation_forms_agree():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b = rng.uniform(0.1, 2.0, 2) ** 2
        g = rng.uniform(-1, 1) * np.sqrt(a * b) * 0.9
        zeta = (b - a) / (2 * g)
        t0 = np.copysign(1.0, zeta) / (abs(zeta) + np.sqrt(1 + zeta * zeta))
        d = b - a
        t1 = (2 * g if d >= 0 else -2 * g) / (abs(d) + np.sqrt(d * d + 4 * g * g))
        assert abs(t0 - t1) <= 4e-16 * max(1.0, abs(t0))
        c, s = 1 / np.sqrt(1 + t1 * t1), t1 / np.sqrt(1 + t1 * t1)
        assert abs((c * c - s * s) * g + c * s * (a - b)) <= 1e-15 * np.sqrt(a * b)   # the rotated columns are orthogonal


def test_preconditioned_jacobi_factorisation_and_sweep_count():
    rng = np.random.default_rng(3)
    N = 32
    A = _graded(N, 2.0, rng)
    eps_rot = max(1e-15, np.sqrt(N) * 2.220446049250313e-16)
    plain, G, V = _sweeps(A, np.eye(N), eps_rot)
    Q0, R = np.linalg.qr(A)
    pre, Gp, Vp = _sweeps(R.T.copy(), Q0, eps_rot)
    assert pre <= 10 and plain >= pre + 5, (plain, pre)
    ref = np.linalg.svd(A, compute_uv=False)
    for Gx, Vx, left_is_accumulator in ((G, V, False), (Gp, Vp, True)):
        sig = np.linalg.norm(Gx, axis=0)
        o = np.argsort(-sig)
        sig, Gn, Vn = sig[o], Gx[:, o] / sig[o], Vx[:, o]
        P, Q = (Vn, Gn) if left_is_accumulator else (Gn, Vn)
        assert np.max(np.abs(sig - ref)) <= 1e-14 * ref[0]
        assert np.linalg.norm(P @ np.diag(sig) @ Q.T - A) <= 1e-13 * np.linalg.norm(A)
        assert np.linalg.norm(Vn.T @ Vn - np.eye(N)) <= 1e-12          # the accumulator is orthogonal in both variants
