"""Run under torchrun (one rank per GPU): row-sharded steps through libdlra.so + NCCL versus the CPU oracle on the
unsharded problem.  Prints 'MULTI_GPU_OK' on rank 0 when every check passed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lowrankintegrators.jl_b200 as lri  # noqa: E402
from oracle import dlra_oracle as O  # noqa: E402
from tests.problems import lowrank_stream, rel_fro  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = []
    for (n, m, r, name) in [(4096, 512, 8, "bug"), (4096, 512, 8, "ksl_primal"), (4096, 512, 8, "ksl_dual"), (6144, 384, 16, "bug"),
                            (4096, 512, 6, "rabug"), (4096, 512, 8, "greedy"), (2050, 130, 5, "bug"),
                            (4096, 512, 8, "greedy2"), (2050, 130, 5, "greedy2"),
                            (16384, 384, 40, "bug"), (16384, 384, 24, "ksl_primal")]:
        A = lowrank_stream(n, m, 2 * r if name != "rabug" else 10, seed=21, eps=0.0 if name == "rabug" else 1e-4)
        snaps = [A(0.04 * k) for k in range(4)]
        X0 = O.truncated_svd(snaps[0], r)
        lo, hi = lri.row_shard(n, world, rank)
        galg, oalg = {
            "bug": (lri.UnconventionalAlgorithm(), O.UnconventionalAlgorithm()),
            "ksl_primal": (lri.ProjectorSplitting(lri.PrimalLieTrotter()), O.ProjectorSplitting(O.PrimalLieTrotter())),
            "ksl_dual": (lri.ProjectorSplitting(lri.DualLieTrotter()), O.ProjectorSplitting(O.DualLieTrotter())),
            "rabug": (lri.RankAdaptiveUnconventionalAlgorithm(1e-6, rmax=16), O.RankAdaptiveUnconventionalAlgorithm(1e-6, rmax=16)),
            "greedy": (lri.GreedyIntegrator(), O.GreedyIntegrator()),
            "greedy2": (lri.GreedyIntegrator(), O.GreedyIntegrator()),
        }[name]
        dsn = [torch.from_numpy(np.ascontiguousarray(s[lo:hi].T)).cuda().t() for s in snaps]
        if name == "greedy2":   # u = U*Z' (greedy_integrator.jl:84-92): U row-sharded, Z replicated
            Z0 = snaps[0].T @ X0.U
            gu0, ou0 = lri.TwoFactorRepresentation(X0.U[lo:hi], Z0), O.TwoFactorRepresentation(X0.U, Z0)
        else:
            gu0, ou0 = lri.SVDLikeRepresentation(X0.U[lo:hi], X0.S, X0.V), X0
        gint = lri.init(lri.MatrixDataProblem(dsn, gu0), galg, 1, comm="torch")
        oint = O.init(O.MatrixDataProblem(snaps, ou0), oalg, 1)
        for k in range(3):
            O.step(oint)
            lri.step(gint)
            gu = gint.u
            parts = [None] * world
            dist.all_gather_object(parts, gu.U)
            Ufull = np.vstack(parts)
            if rank == 0:
                ou = oint.u
                ok_rank = gu.rank == ou.rank
                full = Ufull @ gu.Z.T if name == "greedy2" else Ufull @ gu.S @ gu.V.T
                err = rel_fro(full, ou.full()) if ok_rank else np.inf
                orth = np.linalg.norm(Ufull.T @ Ufull - np.eye(gu.rank))
                if not (ok_rank and err <= 1e-10 and orth < 1e-12):
                    failures.append((n, m, r, name, k, gu.rank, ou.rank, err, orth))
        gint.cache.close()
    if rank == 0:
        print("FAILURES", failures)
        print("MULTI_GPU_OK" if not failures else "MULTI_GPU_FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
