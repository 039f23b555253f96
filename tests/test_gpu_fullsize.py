"""Full BASELINE-size checks: a direct 3-step comparison with the CPU oracle (about 1.3 s per oracle step) and size-independent properties:
exactness — if rank A(t) <= r, KSL and BUG reproduce A(t_k) to round-off (README refs [1],[2] of the reference) — at
configs[1] size n=65536, m=4096, r=16 with device-resident snapshots, for the two-pass and the software-pipelined path,
and pipelined == two-pass on a full-rank perturbed stream."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, M, R = 65536, 4096, 16


@pytest.fixture(scope="module")
def setup():
    import torch
    import lowrankintegrators.jl_b200 as lri
    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    P = torch.rand((N, R), generator=g, device=dev, dtype=torch.float64) * 2 - 1
    Wm = torch.rand((M, R), generator=g, device=dev, dtype=torch.float64) * 2 - 1
    sig = 2.0 ** -(0.5 * torch.arange(R, device=dev, dtype=torch.float64))
    om = torch.linspace(0.5, 2.0, R, device=dev, dtype=torch.float64)

    def coeff(k):
        return sig * torch.cos(om * (0.05 * k) + 0.3)

    snaps = []
    for k in range(4):
        A = lri.empty_colmajor(N, M, dev)
        A.copy_((P * coeff(k)) @ Wm.T)
        snaps.append(A)
    # exact rank-R factors of A_0 from small factorisations (the n x m SVD is never formed)
    Qp, Rp = torch.linalg.qr(P * coeff(0))
    Qw, Rw = torch.linalg.qr(Wm)
    Us, s, Vsh = torch.linalg.svd(Rp @ Rw.T)
    U0, S0, V0 = Qp @ Us, torch.diag(s), Qw @ Vsh.T
    return lri, torch, snaps, (U0, S0, V0)


def _run(lri, snaps, u0, alg, lookahead):
    L = lri._lib
    eng = lri.Engine(N, M, R)
    eng.set_factors(*u0)
    eng.data_init(snaps[0])
    errs = []
    pushed = 0
    for k in range(len(snaps) - 1):
        if pushed == 0:
            eng.data_push(snaps[k + 1]); pushed = 1
        if lookahead and pushed == 1 and k + 2 < len(snaps):
            eng.data_push(snaps[k + 2]); pushed = 2
        pushed -= 1
        if alg == "bug":
            eng.step_bug()
        else:
            eng.step_ksl(L.KSL_PRIMAL if alg == "ksl_primal" else L.KSL_DUAL)
        errs.append(eng.reconstruct_error(snaps[k + 1]))
    fac = eng.get_factors_device()
    eng.close()
    return errs, fac


@pytest.mark.parametrize("alg,lookahead", [("bug", False), ("bug", True), ("ksl_primal", False), ("ksl_dual", False)])
def test_exactness_at_config2_size(setup, alg, lookahead):
    lri, torch, snaps, u0 = setup
    errs, (U, S, V) = _run(lri, snaps, u0, alg, lookahead)
    assert max(errs) < 1e-12, errs
    eye = torch.eye(R, device=U.device, dtype=torch.float64)
    assert float(torch.linalg.norm(U.T @ U - eye)) < 1e-12 and float(torch.linalg.norm(V.T @ V - eye)) < 1e-12


def test_pipelined_equals_two_pass_on_full_rank_stream(setup):
    lri, torch, snaps, u0 = setup
    g = torch.Generator(device=snaps[0].device)
    g.manual_seed(11)
    noisy = []
    for A in snaps:
        B = lri.empty_colmajor(N, M, A.device)
        B.copy_(A)
        B.add_(1e-3 * (torch.rand((N, M), generator=g, device=A.device, dtype=torch.float64) - 0.5))
        noisy.append(B)
    _, (U1, S1, V1) = _run(lri, noisy, u0, "bug", False)
    _, (U2, S2, V2) = _run(lri, noisy, u0, "bug", True)
    # ||U1 S1 V1' - U2 S2 V2'||_F accumulated over column blocks (a Gram-based formula would cancel catastrophically)
    num = den = 0.0
    US1, US2 = U1 @ S1, U2 @ S2
    for j0 in range(0, M, 256):
        Y1 = US1 @ V1[j0:j0 + 256].T
        Y2 = US2 @ V2[j0:j0 + 256].T
        num += float(torch.sum((Y1 - Y2) ** 2))
        den += float(torch.sum(Y1 ** 2))
    rel = (num / den) ** 0.5
    assert rel < 1e-10, rel


def _blockwise_rel(torch, fa, fb):
    (U1, S1, V1), (U2, S2, V2) = fa, fb
    num = den = 0.0
    US1, US2 = U1 @ S1, U2 @ S2
    for j0 in range(0, M, 256):
        Y1 = US1 @ V1[j0:j0 + 256].T
        Y2 = US2 @ V2[j0:j0 + 256].T
        num += float(torch.sum((Y1 - Y2) ** 2))
        den += float(torch.sum(Y2 ** 2))
    return (num / den) ** 0.5


@pytest.mark.parametrize("lookahead", [False, True])
def test_oracle_parity_at_config2_size(setup, lookahead):
    """Direct comparison with the CPU oracle at the FULL configs[1] size (65536 x 4096, r = 16, BUG, unconventional.jl:133-157):
    three steps from the same u0 on the same bytes, two-pass and software-pipelined path, 1e-10 on U*S*V' after every step."""
    from oracle import dlra_oracle as O
    lri, torch, snaps, u0 = setup
    g = torch.Generator(device=snaps[0].device)
    g.manual_seed(23)
    noisy = []
    for A in snaps:
        B = lri.empty_colmajor(N, M, A.device)
        B.copy_(A)
        B.add_(1e-3 * (torch.rand((N, M), generator=g, device=A.device, dtype=torch.float64) - 0.5))
        noisy.append(B)
    host = [np.asfortranarray(B.t().cpu().numpy().T) for B in noisy]   # the same bytes, column-major on the host
    U0, S0, V0 = (x.cpu().numpy() for x in u0)
    oint = O.init(O.MatrixDataProblem(host, O.SVDLikeRepresentation(U0, S0, V0)), O.UnconventionalAlgorithm(), 1)
    L = lri._lib
    eng = lri.Engine(N, M, R)
    eng.set_factors(*u0)
    eng.data_init(noisy[0])
    pushed = 0
    errs = []
    for k in range(len(noisy) - 1):
        if pushed == 0:
            eng.data_push(noisy[k + 1]); pushed = 1
        if lookahead and pushed == 1 and k + 2 < len(noisy):
            eng.data_push(noisy[k + 2]); pushed = 2
        pushed -= 1
        eng.step_bug()
        O.step(oint)
        gfac = eng.get_factors_device()
        ofac = tuple(torch.from_numpy(np.ascontiguousarray(x)).to(gfac[0].device) for x in (oint.u.U, oint.u.S, oint.u.V))
        errs.append(_blockwise_rel(torch, gfac, ofac))
    eng.close()
    assert max(errs) <= 1e-10, errs
