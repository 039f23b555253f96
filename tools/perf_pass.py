"""Quick timing of the data-problem steps at a given shape (device-resident snapshots)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lowrankintegrators.jl_b200 as lri
L = lri._lib

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    r = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    algs = sys.argv[5].split(",") if len(sys.argv) > 5 else ["bug", "ksl", "rabug", "greedy"]
    kinds = sys.argv[6].split(",") if len(sys.argv) > 6 else ["delta", "snapshot"]
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(0)
    snaps = [lri.empty_colmajor(n, m, dev) for _ in range(3)]
    for s in snaps: s.copy_(torch.rand((n, m), generator=g, device=dev, dtype=torch.float64) - 0.5)
    U0 = torch.linalg.qr(torch.randn((n, r), generator=g, device=dev, dtype=torch.float64))[0]
    V0 = torch.linalg.qr(torch.randn((m, r), generator=g, device=dev, dtype=torch.float64))[0]
    S0 = torch.diag(torch.tensor([2.0 ** -i for i in range(r)], device=dev, dtype=torch.float64))
    for alg in algs:
        for kind, kname in ((L.DATA_DELTA, "delta"), (L.DATA_SNAPSHOT, "snapshot")):
            if alg in ("greedy", "greedy2", "normal") and kind == L.DATA_DELTA: continue
            if kname not in kinds: continue
            eng = lri.Engine(n, m, r, rmax=r, rank_adaptive=(alg == "rabug"),
                             aug_basis_first=(alg == "rabug" and os.environ.get("DLRA_AUG") == "1"))
            eng.set_factors(U0, S0, V0)
            eng.data_init(snaps[0])
            look = len(sys.argv) > 7 and sys.argv[7] == "lookahead" and alg == "bug" and kind == L.DATA_SNAPSHOT
            if look: eng.data_push(snaps[1], kind)
            def one(i):
                if alg == "normal":
                    eng.normal_component(snaps[i % 3]); return
                eng.data_push(snaps[(i + 2) % 3] if look else snaps[(i + 1) % 3], kind)
                if alg == "bug": eng.step_bug()
                elif alg == "ksl": eng.step_ksl(L.KSL_PRIMAL)
                elif alg == "rabug": eng.step_rabug(1e-3, r)
                elif alg == "greedy2": eng.step_greedy_two_factor(L.GREEDY_DATA)
                else: eng.step_greedy()
            for i in range(3): one(i)
            eng.sync(); eng.set_profiling(True); eng.stats(reset=True)
            t0 = time.perf_counter()
            for i in range(steps): one(i)
            eng.sync(); dt = (time.perf_counter() - t0) / steps
            st = eng.stats()
            pm = st["pass_ms"] / steps
            gbs = st["pass_bytes"] / (st["pass_ms"] * 1e-3) / 1e9 if st["pass_ms"] > 0 else 0
            print(f"{alg:6s} {kname:8s} n={n} m={m} r={r}: {dt*1e3:8.3f} ms/step  passes {pm:7.3f} ms/step "
                  f"({st['pass_launches']//steps} launches/step, {gbs:7.1f} GB/s)  kernels/step {st['kernel_launches']/steps:.0f}", flush=True)
            eng.close()
main()
