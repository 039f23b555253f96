// The K/L/S contractions over the dense increment ΔA — `u .+= sign*left'*Δy*right`
// (data_integrator.jl:13-16) — as two fused streaming passes per step:
//   pass_KL : K += ΔA·Vf  (n x r, complete rows)   and/or   Lout = ΔAᵀ·Uf  (m x r, reduced over row panels)
//   pass_S  : Sout = Lfᵀ·ΔA·Rf   (p x q core: K-use W = ΔA·Rf, then a tall-skinny Gram product Lfᵀ·W)
// ΔA is either a pre-differenced increment or formed on the fly as A − Aprev (the reference's
// `Δy .= ycurr - yprev`, projector_splitting.jl:119-121) so no n x m temporary is ever written.
// Dispatch: TMA + DMMA kernels (pass_tma.cuh) when the shape/alignment allows, else the generic GEMMs.
#pragma once
#include "engine.cuh"
#include "pass_tma.cuh"
#include "pass_tri.cuh"

namespace dlra {

inline double delta_bytes(const dlra_engine* e, const Delta& d) { return (double)e->n * (double)e->m * 8.0 * (d.Aprev ? 2.0 : 1.0); }

// K (n x r, ldk) += ΔA·Vf  if K != nullptr (this rank's rows);
// Lout (m x r, ldl) = Σ_ranks ΔAᵀ·Uf + Vi·Siᵀ  if Lout != nullptr  (complete: all-reduced over row shards, initial term added)
inline void pass_KL(dlra_engine* e, const Delta& d, int r, const double* Vf, int64_t ldv, const double* Uf, int64_t ldu,
                    double* K, int64_t ldk, double* Lout, int64_t ldl, const double* Vi = nullptr, int64_t ldvi = 0,
                    const double* Si = nullptr, int64_t ldsi = 0) {
    Ctx& cx = e->cx;
    NvtxRange nvtx_pass("dlra:pass_KL");
    if (!(e->flags & DLRA_FORCE_GENERIC) && tma_pass_supported(e->n, e->m, d)) {
        tma_pass_KL(e, d, r, Vf, ldv, Uf, ldu, K, ldk, Lout, ldl, Vi, ldvi, Si, ldsi);   // timed per kernel launch inside
        return;
    }
    if (K) {
        pass_timer_begin(e, delta_bytes(e, d), 1, 2.0 * (double)e->n * (double)e->m * r);
        gemm_nn(cx, e->n, (int)e->m, r, d.A, d.lda, d.Aprev, d.ldap, Vf, ldv, false, K, ldk, 1.0, 1.0);
        pass_timer_end(e);
    }
    if (Lout) {
        e->gws.ensure(gemm_tn_ws(cx, e->n, (int)e->m, r), cx.stream);
        pass_timer_begin(e, delta_bytes(e, d), 2, 2.0 * (double)e->n * (double)e->m * r);
        gemm_tn(cx, e->n, (int)e->m, r, d.A, d.lda, d.Aprev, d.ldap, Uf, ldu, Lout, ldl, 1.0, 0.0, e->gws.p);
        pass_timer_end(e);
        if (ldl == e->m) e->comm.allreduce_sum(Lout, e->m * (int64_t)r, cx);
        else for (int c = 0; c < r; ++c) e->comm.allreduce_sum(Lout + (int64_t)c * ldl, e->m, cx);
        if (Vi) gemm_nn(cx, e->m, r, r, Vi, ldvi, nullptr, 0, Si, ldsi, true, Lout, ldl, 1.0, 1.0);
    }
}

// Sout (p x q, lds) = Lfᵀ·ΔA·Rf  (this rank's rows only);  Lf: n x p, Rf: m x q
inline void pass_S(dlra_engine* e, const Delta& d, int p, int q, const double* Lf, int64_t ldlf, const double* Rf, int64_t ldrf,
                   double* Sout, int64_t lds) {
    Ctx& cx = e->cx;
    NvtxRange nvtx_pass("dlra:pass_S");
    if (!(e->flags & DLRA_FORCE_GENERIC) && tma_pass_supported(e->n, e->m, d)) {
        tma_pass_S(e, d, p, q, Lf, ldlf, Rf, ldrf, Sout, lds);
        return;
    }
    e->nscr.ensure(e->n * (int64_t)q, cx.stream);
    pass_timer_begin(e, delta_bytes(e, d), 1, 2.0 * (double)e->n * (double)e->m * q);
    gemm_nn(cx, e->n, (int)e->m, q, d.A, d.lda, d.Aprev, d.ldap, Rf, ldrf, false, e->nscr.p, e->n, 1.0, 0.0);
    pass_timer_end(e);
    e->gws.ensure(gemm_tn_ws(cx, e->n, p, q), cx.stream);
    gemm_tn(cx, e->n, p, q, Lf, ldlf, nullptr, 0, e->nscr.p, e->n, Sout, lds, 1.0, 0.0, e->gws.p);
}

}  // namespace dlra
