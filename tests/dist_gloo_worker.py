"""world_size-2 CPU (gloo) emulation of the row-sharded step algebra the engine runs over NCCL:
K/Uhat row-sharded, L/S/M partial sums all-reduced, TSQR R-factors all-gathered + redundant small QR.
Checked against the unsharded oracle.  Run under torchrun."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dlra_oracle as O  # noqa: E402
from tests.problems import lowrank_stream, rel_fro  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "lowrankintegrators.jl_b200"))
from distributed import row_shard  # noqa: E402  (pure-python helper, no CUDA needed)


def allreduce(x):
    t = torch.from_numpy(np.ascontiguousarray(x))
    dist.all_reduce(t)
    return t.numpy()


def dist_tsqr(A):
    """local Householder QR -> all-gather R -> redundant QR of the stack -> Q_local * Q_top[block]."""
    world, rank = dist.get_world_size(), dist.get_rank()
    Q1, R1 = np.linalg.qr(A)
    Rs = [torch.zeros(R1.shape, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(Rs, torch.from_numpy(np.ascontiguousarray(R1)))
    Q2, R = np.linalg.qr(np.vstack([r.numpy() for r in Rs]))
    c = R1.shape[0]
    return Q1 @ Q2[rank * c:(rank + 1) * c], R


def sharded_bug_step(U, S, V, dA):
    K = U @ S + dA @ V
    L = V @ S.T + allreduce(dA.T @ U)
    U1, _ = dist_tsqr(K)
    M = allreduce(U1.T @ U)
    V1, _ = np.linalg.qr(L)
    N = V1.T @ V
    S1 = M @ S @ N.T + allreduce(U1.T @ dA @ V1)
    return U1, S1, V1


def sharded_ksl_step(U, S, V, dA):
    K = U @ S + dA @ V
    U1, R = dist_tsqr(K)
    Wm = allreduce(dA.T @ U1)
    St = R - Wm.T @ V
    L = V @ St.T + Wm
    V1, RL = np.linalg.qr(L)
    return U1, RL.T, V1


def sharded_greedy_two_factor_step(U, Z, X):
    """greedy_integrator.jl:84-92 as the engine runs it: Z all-reduced, polar factor via TSQR + SVD of the small R."""
    Z1 = allreduce(X.T @ U)
    Q, R = dist_tsqr(X @ Z1)
    P, _, Qt = np.linalg.svd(R)
    return Q @ (P @ Qt), Z1


def main():
    dist.init_process_group("gloo")
    world, rank = dist.get_world_size(), dist.get_rank()
    n, m, r = 257, 96, 6
    lo, hi = row_shard(n, world, rank)
    spans = [row_shard(n, world, k) for k in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    A = lowrank_stream(n, m, 12, seed=13, eps=1e-4)
    snaps = [A(0.05 * k) for k in range(4)]
    X0 = O.truncated_svd(snaps[0], r)
    ok = True
    for name, fn, oalg in (("bug", sharded_bug_step, O.UnconventionalAlgorithm()),
                           ("ksl", sharded_ksl_step, O.ProjectorSplitting(O.PrimalLieTrotter()))):
        U, S, V = X0.U[lo:hi].copy(), X0.S.copy(), X0.V.copy()
        oint = O.init(O.MatrixDataProblem(snaps, X0), oalg, 1)
        for k in range(3):
            U, S, V = fn(U, S, V, (snaps[k + 1] - snaps[k])[lo:hi])
            O.step(oint)
            parts = [None] * world
            dist.all_gather_object(parts, U)
            err = rel_fro(np.vstack(parts) @ S @ V.T, oint.u.full())
            ok = ok and err < 1e-11
    U = X0.U[lo:hi].copy()
    Z = snaps[0].T @ X0.U
    oint = O.init(O.MatrixDataProblem(snaps, O.TwoFactorRepresentation(X0.U, Z)), O.GreedyIntegrator(), 1)
    for k in range(3):
        U, Z = sharded_greedy_two_factor_step(U, Z, snaps[k + 1][lo:hi])
        O.step(oint)
        parts = [None] * world
        dist.all_gather_object(parts, U)
        ok = ok and rel_fro(np.vstack(parts) @ Z.T, oint.u.full()) < 1e-11
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("GLOO_SHARDED_OK" if int(flag) == 1 else "GLOO_SHARDED_FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
