"""Small workload for compute-sanitizer (racecheck / synccheck / memcheck): a few steps of every data integrator at sizes that
take the TMA/DMMA passes (plain and cluster/multicast), the pipelined BUG pass, the CTA-wide TSQR and the cluster Jacobi SVD.
  compute-sanitizer --tool racecheck python tools/sanitize_driver.py
Under torchrun (WORLD_SIZE > 1) the steps run row-sharded (P2P exchange kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lowrankintegrators.jl_b200 as lri
L = lri._lib

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def run(n, m, r, alg, adaptive=False, steps=3, lookahead=False):
    g = torch.Generator(device=dev); g.manual_seed(3 + rank)
    snaps = [lri.empty_colmajor(n, m, dev) for _ in range(3)]
    for s in snaps: s.copy_(torch.rand((n, m), generator=g, device=dev, dtype=torch.float64) - 0.5)
    gw = torch.Generator(device=dev); gw.manual_seed(11)
    U0 = torch.linalg.qr(torch.randn((n, r), generator=g, device=dev, dtype=torch.float64))[0] / np.sqrt(world)
    V0 = torch.linalg.qr(torch.randn((m, r), generator=gw, device=dev, dtype=torch.float64))[0]
    S0 = torch.diag(2.0 ** -torch.arange(r, device=dev, dtype=torch.float64))
    eng = lri.Engine(n, m, r, rmax=r, rank_adaptive=adaptive, device=local)
    if world > 1: lri.attach_engine(eng)
    eng.set_factors(U0, S0, V0)
    eng.data_init(snaps[0])
    if lookahead: eng.data_push(snaps[1])
    for i in range(steps):
        eng.data_push(snaps[(i + 2) % 3] if lookahead else snaps[(i + 1) % 3])
        if alg == "bug": eng.step_bug()
        elif alg == "ksl": eng.step_ksl(L.KSL_PRIMAL)
        elif alg == "ksl_dual": eng.step_ksl(L.KSL_DUAL)
        elif alg == "rabug": eng.step_rabug(1e-3, r)
        else: eng.step_greedy()
    _, S, _ = eng.get_factors()
    assert np.isfinite(S).all()
    eng.close()
    if rank == 0: print(f"ok {alg} n={n} m={m} r={r} lookahead={lookahead}", flush=True)


run(4096, 256, 16, "bug", lookahead=True)      # tri_pass, CTA TSQR, l_finalize
run(4096, 256, 16, "ksl")
run(4096, 256, 16, "ksl_dual")
run(4096, 256, 40, "bug")                      # cluster/multicast pass, BCGS2
run(2048, 256, 16, "rabug", adaptive=True)     # Jacobi (one CTA), 32-wide core pass
run(2048, 320, 72, "rabug", adaptive=True)     # cluster Jacobi (2r = 144 -> 4 CTAs)
run(4096, 256, 8, "greedy")
if world > 1:
    dist.destroy_process_group()
print("SANITIZE_DRIVER_DONE")
