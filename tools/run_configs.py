"""Runs the BASELINE.json configs that are not the bench headline at (or near) full size on one B200 and prints
ms/step, so that memory footprint and scaling cliffs are visible.  Usage: python tools/run_configs.py [cfg1,cfg3,cfg4,cfg5]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import scipy.sparse as sp
import lowrankintegrators.jl_b200 as lri
L = lri._lib
dev = torch.device("cuda", 0)


def csr_dev(M):
    M = sp.csr_matrix(M)
    return (torch.from_numpy(M.indptr.astype(np.int64)).cuda(), torch.from_numpy(M.indices.astype(np.int32)).cuda(),
            torch.from_numpy(M.data.astype(np.float64)).cuda(), M.shape)


def periodic_ops(n, nu=0.005, length=np.pi):
    dx = length / n
    i = np.arange(n)
    lap = sp.csr_matrix((np.r_[np.full(n, nu / dx ** 2), np.full(n, -2 * nu / dx ** 2), np.full(n, nu / dx ** 2)],
                         (np.r_[i, i, i], np.r_[(i - 1) % n, i, (i + 1) % n])), shape=(n, n))
    grad = sp.csr_matrix((np.r_[np.full(n, -0.5 / dx), np.full(n, 0.5 / dx)], (np.r_[i, i], np.r_[(i - 1) % n, (i + 1) % n])), shape=(n, n))
    return lap, grad


def orth(n, r, g):
    return torch.linalg.qr(torch.randn((n, r), generator=g, device=dev, dtype=torch.float64))[0]


def timed(eng, fn, steps, warm=1):
    for _ in range(warm): fn()
    eng.sync(); st0 = eng.stats(reset=True)
    t0 = time.perf_counter()
    for _ in range(steps): fn()
    eng.sync()
    dt = (time.perf_counter() - t0) / steps
    st = eng.stats()
    return dt * 1e3, st["kernel_launches"] / steps


def cfg1():
    # generic_matrix-style MatrixDEProblem n=m=1000, rank 5, KSL primal, dt=1e-2 (dense operators), Tsit5 default
    n, r = 1000, 5
    g = torch.Generator(device=dev); g.manual_seed(0)
    W1 = torch.triu(torch.randn((n, n), generator=g, device=dev, dtype=torch.float64), 1); W1 = W1 - W1.T
    W2 = torch.triu(torch.randn((n, n), generator=g, device=dev, dtype=torch.float64), 1); W2 = W2 - W2.T
    A = lri.colmajor_device(W1 + torch.eye(n, device=dev, dtype=torch.float64)); B = lri.colmajor_device(W2.T.contiguous())
    eng = lri.Engine(n, n, r)
    eng.set_factors(orth(n, r, g), torch.diag(2.0 ** -torch.arange(1, r + 1, device=dev, dtype=torch.float64)), orth(n, r, g))
    eng.rhs_set(A=A, B=B)
    t = [0.0]
    def one():
        eng.step_ksl(L.KSL_PRIMAL, t[0], 1e-2); t[0] += 1e-2
    ms, kl = timed(eng, one, 20)
    print(f"cfg1 DE generic matrix n=m=1000 r=5 KSL primal (adaptive Tsit5): {ms:.3f} ms/step, {kl:.0f} kernels/step", flush=True)
    eng.close()


def cfg3():
    # Burgers UQ n=8192 grid x m=16384 samples, r=32, KSL primal, column-wise nonlinear F on device
    n, mm, r = 8192, 128, 32
    m = mm * mm
    g = torch.Generator(device=dev); g.manual_seed(0)
    lap, grad = periodic_ops(n)
    eng = lri.Engine(n, m, r)
    U0, S0, V0 = orth(n, r, g), torch.diag(2.0 ** -torch.arange(r, device=dev, dtype=torch.float64)), orth(m, r, g)
    eng.rhs_set(A=csr_dev(lap), D1=csr_dev(grad), D2=1.0, c_had=-1.0)
    # the discrete Laplacian has |lambda|max = 4 nu/dx^2 = 1.4e5: dt = 1e-5 keeps one RK4 stage set per flow stable
    # (timing a diverging run would be meaningless); the adaptive sub-stepper picks its own sub-steps below dt
    dt = 1e-5
    for sub, nm in ((L.ODE_RK4, "rk4 x1"), (L.ODE_TSIT5, "adaptive Tsit5")):
        eng.set_factors(U0, S0, V0)
        for f in (L.FLOW_K, L.FLOW_S, L.FLOW_L): eng.set_substepper(f, sub, 1)
        t = [0.0]
        def one():
            eng.step_ksl(L.KSL_PRIMAL, t[0], dt); t[0] += dt
        ms, kl = timed(eng, one, 5)
        _, S, _ = eng.get_factors()
        assert np.isfinite(S).all(), "cfg3 diverged"
        print(f"cfg3 Burgers n={n} m={m} r={r} KSL primal dt={dt} ({nm}): {ms:.3f} ms/step, {kl:.0f} kernels/step, "
              f"|S|={np.linalg.norm(S):.4f}", flush=True)
    eng.close()


def cfg4():
    # rank-adaptive BUG on F(X) = A X + X B' + G H' (stencil operators, low-rank forcing), n=m=262144, r 8 -> 128
    n = m = 262144
    g = torch.Generator(device=dev); g.manual_seed(0)
    A = sum(periodic_ops(n, nu=0.02)); B = sum(periodic_ops(m, nu=0.03))
    q = 128
    G = orth(n, q, g); H = orth(m, q, g) * (0.7 ** torch.arange(q, device=dev, dtype=torch.float64))
    eng = lri.Engine(n, m, 8, rmax=128, rank_adaptive=True, aug_basis_first=os.environ.get('CFG4_AUG') == '1')
    eng.set_factors(orth(n, 8, g), torch.diag(2.0 ** -torch.arange(8, device=dev, dtype=torch.float64)), orth(m, 8, g))
    eng.rhs_set(A=csr_dev(A), B=csr_dev(B), G=lri.colmajor_device(G), H=lri.colmajor_device(H))
    for f in (L.FLOW_K, L.FLOW_S, L.FLOW_L): eng.set_substepper(f, L.ODE_RK4, 1)
    t = 0.0
    for k in range(int(os.environ.get('CFG4_STEPS', '8'))):
        eng.sync(); t0 = time.perf_counter()
        rn, ch = eng.step_rabug(1e-8, 128, t, 1e-6)
        eng.sync(); ms = (time.perf_counter() - t0) * 1e3
        t += 1e-6
        print(f"cfg4 Lyapunov n=m={n} RA-BUG step {k}: rank -> {rn} ({ms:.1f} ms)", flush=True)
        if os.environ.get("DLRA_PHASES"): eng.stats()
    eng.close()


def cfg5():
    # large-scale data compression, r=64; one GPU holds a 2^20-row shard of the pre-differenced stream (32 GiB)
    n, m, r = 1 << 20, 4096, 64
    g = torch.Generator(device=dev); g.manual_seed(0)
    dA = lri.empty_colmajor(n, m, dev)
    for j0 in range(0, m, 512):
        dA[:, j0:j0 + 512] = torch.rand((n, 512), generator=g, device=dev, dtype=torch.float64) - 0.5
    eng = lri.Engine(n, m, r)
    eng.set_factors(orth(n, r, g), torch.diag(2.0 ** -(0.25 * torch.arange(r, device=dev, dtype=torch.float64))), orth(m, r, g))
    def one():
        eng.data_push(dA, L.DATA_DELTA); eng.step_bug()
    eng.set_profiling(True)
    ms, kl = timed(eng, one, 3)
    br = eng.pass_breakdown()
    print(f"cfg5 shard n=2^20 m=4096 r=64 BUG (delta stream): {ms:.2f} ms/step, {kl:.0f} kernels/step; "
          + ", ".join(f"{k}: {v['ms']/max(v['launches'],1):.2f} ms x{v['launches']//3}/step {v['flops']/max(v['ms'],1e-9)/1e9:.1f} TF" for k, v in br.items() if v['launches']), flush=True)
    eng.close()


if __name__ == "__main__":
    which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg1", "cfg3", "cfg4", "cfg5"]
    for w in which:
        globals()[w]()
        torch.cuda.empty_cache()
