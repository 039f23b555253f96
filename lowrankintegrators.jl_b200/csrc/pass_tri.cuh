// Software-pipelined BUG pass: the core pass of step k and the K/L pass of step k+1 share ONE sweep over the snapshots.
//
// With snapshots A0 = A(t_k), A1 = A(t_k+dt), A2 = A(t_k+2dt) resident and the NEW bases (U1, V1) of step k known,
//     W     = (A1 − A0)·V1        -> core increment of step k      (U1ᵀ·W, unconventional.jl:154-155)
//     K'    = (A2 − A1)·V1        -> K-step of step k+1            (its U0, V0 ARE U1, V1; unconventional.jl:137-139)
//     Lpart = (A2 − A1)ᵀ·U1       -> L-step of step k+1            (unconventional.jl:145-147)
// are all linear in the data and independent of S, so they can be formed together: every snapshot is then read from HBM
// three times over its lifetime instead of four (6·8nm -> ... 3·8nm bytes per step instead of 4·8nm).  The arithmetic is
// identical to the two-kernel path (same operands, same fragment order), only the order of independent work changes.
//
// Same machinery as pass_tma.cuh (TMA 16-row SWIZZLE_128B boxes, mbarrier ring, DMMA.8x8x4, reducer warp, setmaxnreg);
// differences: three tiles per stage, the U sub-tile travels with every stage (L2-resident, evict_last) instead of a
// resident panel so that three 60 KB stages fit, and the K-use keeps two accumulator sets.
#pragma once
#include "pass_tma.cuh"

namespace dlra {

struct TriParams {
    int64_t n, m;
    int rc, nsub, npanels, ntj;
    double* W; int64_t ldw;       // = (A1 − A0)·Vf
    double* K; int64_t ldk;       // = (A2 − A1)·Vf
    double* Lpart; int64_t ldlp;  // per-CTA partial of (A2 − A1)ᵀ·Uf
};

template <int RT>
struct TriSmem {
    static constexpr int VBYTES = PT_TJ * RT * 8;
    static constexpr int UBOX_BYTES = 16 * RT * 8;
    static constexpr int STAGE_BYTES = ((3 * PT_TBYTES + 4 * UBOX_BYTES + VBYTES + 1023) / 1024) * 1024;
    static constexpr int LRED_BYTES = ((PT_CONSUMERS * PT_LRED_LD * RT * 8 + 1023) / 1024) * 1024;
    static constexpr int NST = (225 * 1024 - LRED_BYTES - 1024) / STAGE_BYTES;
    static constexpr int TOTAL = NST * STAGE_BYTES + LRED_BYTES + 1024 + 256;
};

template <int RT>
__global__ void __launch_bounds__(PT_THREADS, 1)
tri_pass_kernel(const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map1,
                const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap mapU,
                const __grid_constant__ CUtensorMap mapV, const TriParams prm) {
    using SM = TriSmem<RT>;
    constexpr int NST = SM::NST;
    constexpr int NB = RT / 8;
    constexpr int LD = PT_LRED_LD;
    constexpr int OFF_U = 3 * PT_TBYTES;
    constexpr int OFF_V = OFF_U + 4 * SM::UBOX_BYTES;
    static_assert(NST >= 2, "tri pass needs at least two stages");
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* stages = base;
    double* lred = reinterpret_cast<double*>(stages + (size_t)NST * SM::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(lred) + SM::LRED_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + NST;
    uint64_t* lfull = bars + 2 * NST;
    uint64_t* lfree = bars + 2 * NST + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], PT_CONSUMERS); }
        mbar_init(lfull, PT_CONSUMERS);
        mbar_init(lfree, PT_NRED);
        mbar_fence_init();
    }
    __syncthreads();
    const int nsub = prm.nsub;

    if (warp >= PT_CONSUMERS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        if (warp == PT_CONSUMERS) {
            if (lane == 0) {
                prefetch_tmap(&map2); prefetch_tmap(&map1); prefetch_tmap(&map0); prefetch_tmap(&mapU); prefetch_tmap(&mapV);
                const uint64_t pol_stream = policy_evict_first();
                const uint64_t pol_keep = policy_evict_last();
                int stage = 0; uint32_t phase = 0;
                for (int panel = blockIdx.x; panel < prm.npanels; panel += gridDim.x) {
                    const int row0 = panel * nsub * PT_SI;
                    for (int jt = 0; jt < prm.ntj; ++jt) {
                        for (int s = 0; s < nsub; ++s) {
                            mbar_wait(&empty[stage], phase ^ 1);
                            uint32_t bytes = 3 * PT_TBYTES + 4 * SM::UBOX_BYTES + (s == 0 ? SM::VBYTES : 0);
                            mbar_expect_tx(&full[stage], bytes);
                            unsigned char* sb = stages + (size_t)stage * SM::STAGE_BYTES;
                            const int r = row0 + s * PT_SI;
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                tma_load_2d(sb + b * PT_BOXBYTES, &map2, r + 16 * b, jt * PT_TJ, &full[stage], pol_stream);
                                tma_load_2d(sb + PT_TBYTES + b * PT_BOXBYTES, &map1, r + 16 * b, jt * PT_TJ, &full[stage], pol_stream);
                                tma_load_2d(sb + 2 * PT_TBYTES + b * PT_BOXBYTES, &map0, r + 16 * b, jt * PT_TJ, &full[stage], pol_stream);
                                tma_load_2d(sb + OFF_U + b * SM::UBOX_BYTES, &mapU, r + 16 * b, 0, &full[stage], pol_keep);
                            }
                            if (s == 0) {
#pragma unroll
                                for (int b = 0; b < 2; ++b)
                                    tma_load_2d(sb + OFF_V + b * (16 * RT * 8), &mapV, jt * PT_TJ + 16 * b, 0, &full[stage], pol_keep);
                            }
                            if (++stage == NST) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
            return;
        }
        {   // three reducer warps: warp wr owns the factor columns i = wr, wr + 3, ... (see pass_tma.cuh)
            const int wr = warp - PT_CONSUMERS - 1;
            constexpr int NI = (RT + PT_NRED - 1) / PT_NRED;
            uint32_t ph = 0;
            double* lp = prm.Lpart + (size_t)blockIdx.x * prm.ldlp * RT;
            for (int panel = blockIdx.x; panel < prm.npanels; panel += gridDim.x) {
                const bool first = (panel == (int)blockIdx.x);
                for (int jt = 0; jt < prm.ntj; ++jt) {
                    const int64_t col = (int64_t)jt * PT_TJ + lane;
                    const bool okc = col < prm.m;
                    double prev[NI];
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        const int i = wr + q * PT_NRED;
                        prev[q] = (!first && okc && i < RT) ? __ldcg(lp + col + (int64_t)i * prm.ldlp) : 0.0;
                    }
                    mbar_wait(lfull, ph);
#pragma unroll
                    for (int q = 0; q < NI; ++q) {
                        const int i = wr + q * PT_NRED;
                        if (i < RT) {
                            double sum = prev[q];
#pragma unroll
                            for (int w = 0; w < PT_CONSUMERS; ++w) sum += lred[(size_t)w * LD * RT + i * LD + lane];
                            if (okc) __stcg(lp + col + (int64_t)i * prm.ldlp, sum);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(lfree);
                    ph ^= 1;
                }
            }
            return;
        }
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");

    const int g = lane >> 2, k = lane & 3;
    const int wbox = warp >> 1, wblk = warp & 1;
    const int prow = (g & 1) + ((g >> 1) & 1) * 8 + (g >> 2) * 2 + wblk * 4;
    const uint32_t offKe = k * 128 + (((prow >> 1) ^ k) << 4) + (prow & 1) * 8;
    const uint32_t offKo = k * 128 + (((prow >> 1) ^ (4 + k)) << 4) + (prow & 1) * 8;
    uint32_t offL[2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const int lrow = (k & 1) + ((k >> 1) & 1) * 8 + 2 * kk + 4 * wblk;
        offL[kk] = g * 128 + (((lrow >> 1) ^ g) << 4) + (lrow & 1) * 8;
    }
    uint32_t offV[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) offV[q] = g * 128 + ((((2 * q) + (k >> 1)) ^ g) << 4) + (k & 1) * 8;
    const int crow = wbox * 16 + prow;

    int stage = 0; uint32_t phase = 0; uint32_t lfree_ph = 1;
    for (int panel = blockIdx.x; panel < prm.npanels; panel += gridDim.x) {
        const int64_t row0 = (int64_t)panel * nsub * PT_SI;
        double wacc[PT_NSUB_MAX][NB][2], kacc[PT_NSUB_MAX][NB][2];
#pragma unroll
        for (int s = 0; s < PT_NSUB_MAX; ++s)
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) { wacc[s][nb][0] = 0.0; wacc[s][nb][1] = 0.0; kacc[s][nb][0] = 0.0; kacc[s][nb][1] = 0.0; }
        for (int jt = 0; jt < prm.ntj; ++jt) {
            double lacc[4][NB][2];
            double vf[8][NB];
#pragma unroll
            for (int cb = 0; cb < 4; ++cb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { lacc[cb][nb][0] = 0.0; lacc[cb][nb][1] = 0.0; }
#pragma unroll
            for (int s = 0; s < PT_NSUB_MAX; ++s) {
                if (s < nsub) {
                    mbar_wait(&full[stage], phase);
                    const unsigned char* sb = stages + (size_t)stage * SM::STAGE_BYTES;
                    const unsigned char* tb = sb + wbox * PT_BOXBYTES;   // tile of A2; A1 at +PT_TBYTES, A0 at +2*PT_TBYTES
                    if (s == 0) {
                        const unsigned char* vt = sb + OFF_V;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb)
                                vf[ks][nb] = *reinterpret_cast<const double*>(vt + (ks >> 2) * (16 * RT * 8) + nb * 1024 + offV[ks & 3]);
                    }
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t off = ((ks & 1) ? offKo : offKe) + ks * 512;
                        const double a2 = *reinterpret_cast<const double*>(tb + off);
                        const double a1 = *reinterpret_cast<const double*>(tb + PT_TBYTES + off);
                        const double a0 = *reinterpret_cast<const double*>(tb + 2 * PT_TBYTES + off);
                        const double dk = a2 - a1, dw = a1 - a0;
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
                            dmma884(kacc[s][nb][0], kacc[s][nb][1], dk, vf[ks][nb]);
                            dmma884(wacc[s][nb][0], wacc[s][nb][1], dw, vf[ks][nb]);
                        }
                    }
                    {
                        const unsigned char* ub = sb + OFF_U + (size_t)wbox * SM::UBOX_BYTES;
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            double uf[NB];
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) uf[nb] = *reinterpret_cast<const double*>(ub + offL[kk] + nb * 1024);
#pragma unroll
                            for (int cb = 0; cb < 4; ++cb) {
                                const uint32_t off = offL[kk] + cb * 1024;
                                const double a = *reinterpret_cast<const double*>(tb + off) - *reinterpret_cast<const double*>(tb + PT_TBYTES + off);
#pragma unroll
                                for (int nb = 0; nb < NB; ++nb) dmma884(lacc[cb][nb][0], lacc[cb][nb][1], a, uf[nb]);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
            mbar_wait(lfree, lfree_ph);
            lfree_ph ^= 1;
            double* mine = lred + (size_t)warp * LD * RT;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    mine[(8 * nb + 2 * k) * LD + 8 * cb + g] = lacc[cb][nb][0];
                    mine[(8 * nb + 2 * k + 1) * LD + 8 * cb + g] = lacc[cb][nb][1];
                }
            __syncwarp();
            if (lane == 0) mbar_arrive(lfull);
        }
#pragma unroll
        for (int s = 0; s < PT_NSUB_MAX; ++s) {
            if (s < nsub) {
                const int64_t row = row0 + s * PT_SI + crow;
                if (row < prm.n) {
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int c = 8 * nb + 2 * k + e;
                            if (c < prm.rc) {
                                prm.K[row + (int64_t)c * prm.ldk] = kacc[s][nb][e];   // single chunk: every entry is written exactly once
                                prm.W[row + (int64_t)c * prm.ldw] = wacc[s][nb][e];
                            }
                        }
                }
            }
        }
    }
}

// Launch for one chunk of <= 16 factor columns.  W and K are overwritten; the per-CTA partials of L land in Lpart.
inline void tri_pass_launch(dlra_engine* e, const double* A2, int64_t ld2, const double* A1, int64_t ld1, const double* A0, int64_t ld0,
                            int rc, const double* Vf, int64_t ldv, const double* Uf, int64_t ldu, double* W, int64_t ldw, double* K,
                            int64_t ldk, double* Lpart, int64_t ldlp, int nsub, int npanels) {
    NvtxRange nvtx_pass("dlra:tri_pass");
    using SM = TriSmem<16>;
    auto kern = tri_pass_kernel<16>;
    static unsigned long long attr_devs = 0;
    if (first_use_on_this_device(attr_devs))
        DLRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    CUtensorMap m2 = make_map_2d(A2, e->n, e->m, ld2, 16, PT_TJ, true);
    CUtensorMap m1 = make_map_2d(A1, e->n, e->m, ld1, 16, PT_TJ, true);
    CUtensorMap m0 = make_map_2d(A0, e->n, e->m, ld0, 16, PT_TJ, true);
    CUtensorMap mU = make_map_2d(Uf, e->n, rc, ldu, 16, 16, true);
    CUtensorMap mV = make_map_2d(Vf, e->m, rc, ldv, 16, 16, true);
    TriParams prm;
    prm.n = e->n; prm.m = e->m; prm.rc = rc; prm.nsub = nsub; prm.npanels = npanels; prm.ntj = (int)cdiv(e->m, PT_TJ);
    prm.W = W; prm.ldw = ldw; prm.K = K; prm.ldk = ldk; prm.Lpart = Lpart; prm.ldlp = ldlp;
    const int grid = std::min(npanels, e->cx.num_sms);
    pass_timer_begin(e, 3.0 * (double)e->n * (double)e->m * 8.0, 3, 6.0 * (double)e->n * (double)e->m * rc);
    kern<<<grid, PT_THREADS, SM::TOTAL, e->cx.stream>>>(m2, m1, m0, mU, mV, prm);
    pass_timer_end(e);
    e->cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

}  // namespace dlra
