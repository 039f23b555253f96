// Communication-avoiding tall-skinny Householder QR (TSQR) with explicit thin Q.
// Replaces `qr!(US)` + `Matrix(QRK.Q)` (LAPACK dgeqrt + dgemqrt) at projector_splitting.jl:137-138,149-151,
// 174-175,186-188; unconventional.jl:141-142,149-152; rank_adaptive_unconventional.jl:202-205,213-216.
//
// Every tree node is a Householder factorisation => unconditionally stable, rank-deficient panels included
// (the DLRA K/L matrices routinely have condition numbers ~1/tol, which rules out Gram/Cholesky-QR variants).
//   tsqr_reg_kernel : a CTA owns 1024 rows; each of its 8 warps keeps a 128 x CP panel ENTIRELY IN REGISTERS
//                     (lane l holds rows l, l+32, l+64, l+96), factors it with warp-shuffle reductions
//                     (one batched butterfly per column gives the norm and all trailing dot products), forms
//                     its explicit Q in place (dorg2r recurrence), and the 8 stacked R factors are factored the
//                     same way by warp 0; the explicit Q's are then chained so the CTA writes an orthonormal
//                     1024 x C block plus ONE CP x CP R factor.
//   recursion       : the stacked R factors (nb*CP x CP) go through the same kernel until one CTA suffices;
//                     with row sharding the per-GPU R's are all-gathered (NCCL) and every rank redundantly
//                     factors the G*CP x CP stack.
//   apply_blocks    : Q_block <- Q_block * X_block with the CP x CP blocks of the upper level's Q.
// Columns wider than 16 use block classical Gram-Schmidt with re-orthogonalisation (BCGS2) around 16-column panels.
#pragma once
#include "common.cuh"
#include "comm.cuh"
#include "small_ops.cuh"

namespace dlra {

constexpr int TSQR_NW = 8;      // warps per CTA (4 gives the same step time with one more tree level)
constexpr int TSQR_RPL0 = 4;    // rows per lane at level 0 (128-row panels)
constexpr int TSQR_BR = TSQR_NW * 32 * TSQR_RPL0;  // 1024 rows per CTA
constexpr int TSQR_MAXC = 16;

// Transposing warp reduction: N (power of two) per-lane partial sums -> the warp total of value i ends up on the lanes
// with (lane >> log2(32/N)) == i.  Costs N-1 + log2(32/N) shuffles instead of 5*N for N separate butterflies.
template <int N, int OFF>
struct TReduce {
    static __device__ __forceinline__ double run(double (&e)[N], int lane) {
        double f[N / 2];
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const double keep = up ? e[i + N / 2] : e[i];
            const double send = up ? e[i] : e[i + N / 2];
            f[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        return TReduce<N / 2, OFF / 2>::run(f, lane);
    }
};
template <int OFF>
struct TReduce<1, OFF> {
    static __device__ __forceinline__ double run(double (&e)[1], int) {
        double h = e[0];
#pragma unroll
        for (int o = OFF; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
        return h;
    }
};

// Householder QR of a (32*RPL) x CP register panel: lane l holds rows l + 32 q.  The loop over columns is a REAL loop
// (compact code: a fully unrolled factorisation is instruction-fetch bound) that handles two pivot columns per trip and
// then rotates the register columns by two, so the pivots are always positions 0 and 1.  The finished column j (R above/on
// the diagonal, the Householder vector below it, unit diagonal implicit) is parked in shared memory `vs` ([CP][vld],
// lane <-> row), tau[] as LAPACK dlarfg.  One transposing reduction per column yields the norm and all trailing dots.
// `bc`: 2*CP doubles of warp-private shared memory: the reduced dot products and the pivot row are BROADCAST through it
// (a few LDS per lane) instead of two shuffles per trailing column — the shuffle unit was the busiest pipe of this routine.
template <int CP, int RPL>
__device__ __forceinline__ void reg_panel_qr(double (&a)[RPL][CP], double* vs, int vld, double* tau_s, int lane, double* bc) {
    constexpr int SH = (CP == 16) ? 1 : 2;   // owner lane of reduced value i is i << SH
#pragma unroll 1
    for (int j0 = 0; j0 < CP; j0 += 2) {
        static_for<0, 2>([&](auto PP) {
            constexpr int P = decltype(PP)::value;
            const int j = j0 + P;
            double xm[RPL];
#pragma unroll
            for (int q = 0; q < RPL; ++q) xm[q] = ((q > 0) || (lane > j)) ? a[q][P] : 0.0;
            double e[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                double s = 0.0;
                if (c + P < CP) {
#pragma unroll
                    for (int q = 0; q < RPL; ++q) s = fma(xm[q], a[q][c + P], s);
                }
                e[c] = s;
            }
            const double h = TReduce<CP, 16>::run(e, lane);
            if ((lane & ((1 << SH) - 1)) == 0) bc[lane >> SH] = h;          // value i lives on lanes i << SH ..
            if (lane == j) {
#pragma unroll
                for (int c = P; c < CP; ++c) bc[CP + c] = a[0][c];           // pivot row of the trailing block
            }
            __syncwarp();
            const double alpha = bc[CP + P];
            const double ss = bc[0];
            double t = 0.0, scale = 0.0, beta = alpha;
            if (ss > 0.0) {
                // dlarfg with one rsqrt and one reciprocal on the dependent chain: beta = -sign(alpha)*||x||,
                // tau = (beta - alpha)/beta = 1 + |alpha|/||x||, scale = 1/(alpha - beta)
                const double nrm2 = fma(alpha, alpha, ss);
                const double rn = rsqrt(nrm2);
                beta = -copysign(nrm2 * rn, alpha);
                t = fma(fabs(alpha), rn, 1.0);
                scale = 1.0 / (alpha - beta);
            }
            if (lane == 0) tau_s[j] = t;
            double vm[RPL];
#pragma unroll
            for (int q = 0; q < RPL; ++q) {
                vm[q] = xm[q] * scale;
                a[q][P] = ((q > 0) || (lane > j)) ? vm[q] : a[q][P];
            }
            if (lane == j) { a[0][P] = beta; vm[0] = 1.0; }
#pragma unroll
            for (int q = 0; q < RPL; ++q) vs[j * vld + lane + 32 * q] = a[q][P];
            // apply H_j to the trailing columns
#pragma unroll
            for (int c = P + 1; c < CP; ++c) {
                const double arow = bc[CP + c];
                const double ec = bc[c - P];
                const double w = (arow + ec * scale) * t;
#pragma unroll
                for (int q = 0; q < RPL; ++q) a[q][c] = fma(-w, vm[q], a[q][c]);
            }
            __syncwarp();   // bc is rewritten by the next column
        });
        // rotate two positions to the left
#pragma unroll
        for (int c = 2; c < CP; ++c)
#pragma unroll
            for (int q = 0; q < RPL; ++q) a[q][c - 2] = a[q][c];
#pragma unroll
        for (int q = 0; q < RPL; ++q) { a[q][CP - 2] = 0.0; a[q][CP - 1] = 0.0; }
    }
    __syncwarp();
}

// Explicit thin Q (32*RPL x CP) into registers from the vectors parked by reg_panel_qr (LAPACK dorg2r recurrence, looped,
// two columns per trip with rotating register columns: position c holds Q column j_low + c).
template <int CP, int RPL>
__device__ __forceinline__ void reg_panel_formq(double (&a)[RPL][CP], const double* vs, int vld, const double* tau_s, int lane, double* bc) {
    constexpr int SH = (CP == 16) ? 1 : 2;
#pragma unroll
    for (int c = 0; c < CP; ++c)
#pragma unroll
        for (int q = 0; q < RPL; ++q) a[q][c] = 0.0;
#pragma unroll 1
    for (int j0 = CP - 2; j0 >= 0; j0 -= 2) {
        // rotate two positions to the right: room for columns j0 (position 0) and j0+1 (position 1)
#pragma unroll
        for (int c = CP - 1; c >= 2; --c)
#pragma unroll
            for (int q = 0; q < RPL; ++q) a[q][c] = a[q][c - 2];
        static_for<0, 2>([&](auto PP) {
            constexpr int P = 1 - decltype(PP)::value;   // position 1 (column j0+1) first, then position 0 (column j0)
            const int j = j0 + P;
            const double t = tau_s[j];
            double vm[RPL];
#pragma unroll
            for (int q = 0; q < RPL; ++q) vm[q] = ((q > 0) || (lane > j)) ? vs[j * vld + lane + 32 * q] : 0.0;
            double e[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                double s = 0.0;
                if (c + P + 1 < CP) {
#pragma unroll
                    for (int q = 0; q < RPL; ++q) s = fma(vm[q], a[q][c + P + 1], s);
                }
                e[c] = s;
            }
            const double h = TReduce<CP, 16>::run(e, lane);
            if ((lane & ((1 << SH) - 1)) == 0) bc[lane >> SH] = h;
            __syncwarp();
            if (lane == j) vm[0] = 1.0;    // row j of the finished columns is still 0: the same FMA writes -w there
#pragma unroll
            for (int c = P + 1; c < CP; ++c) {
                const double w = bc[c - P - 1] * t;
#pragma unroll
                for (int q = 0; q < RPL; ++q) a[q][c] = fma(-w, vm[q], a[q][c]);
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < RPL; ++q) a[q][P] = -t * vm[q];
            if (lane == j) a[0][P] = 1.0 - t;
        });
    }
}

template <int CP>
struct TsqrSmem {
    static constexpr int SROWS = TSQR_NW * CP;                 // rows of the stacked R panel
    static constexpr int PROWS = 32 * TSQR_RPL0;               // rows of a level-0 panel
    static constexpr size_t STACK = (size_t)CP * (SROWS + 1);  // doubles
    static constexpr size_t PARK = (size_t)TSQR_NW * CP * PROWS;
    static constexpr size_t BCAST = (size_t)TSQR_NW * 2 * CP;   // per-warp broadcast scratch of the panel routines
    static constexpr size_t BYTES = (2 * STACK + PARK + (size_t)TSQR_NW * CP + BCAST) * sizeof(double);
};

// warp 0: factor the NW*CP x CP stack of R factors (in shared memory), write R to global and leave the explicit Q in place.
template <int CP>
__device__ __noinline__ void tsqr_level1(double* stack, double* stack2, double* taus, double* Rout, int64_t ldr, int lane, double* bc) {
    constexpr int RPL1 = TSQR_NW * CP / 32;
    constexpr int SLD = TSQR_NW * CP + 1;
    double b[RPL1][CP];
#pragma unroll
    for (int c = 0; c < CP; ++c)
#pragma unroll
        for (int q = 0; q < RPL1; ++q) b[q][c] = stack[c * SLD + lane + 32 * q];
    reg_panel_qr<CP, RPL1>(b, stack2, SLD, taus, lane, bc);
    if (lane < CP) {
#pragma unroll
        for (int c = 0; c < CP; ++c) Rout[lane + (int64_t)c * ldr] = (lane <= c) ? stack2[c * SLD + lane] : 0.0;
    }
    reg_panel_formq<CP, RPL1>(b, stack2, SLD, taus, lane, bc);
#pragma unroll
    for (int c = 0; c < CP; ++c)
#pragma unroll
        for (int q = 0; q < RPL1; ++q) stack[c * SLD + lane + 32 * q] = b[q][c];
}

// optional rank-k update applied while the panel is loaded:  A + Ua*Sa  (the `mul!(US, u.U, u.S)` of the K-step,
// unconventional.jl:137, folded into the factorisation so K is not re-read and re-written by a separate GEMM)
struct TsqrAdd {
    const double* U = nullptr; int64_t ldu = 0;   // rows x k
    const double* S = nullptr; int64_t lds = 0;   // k x C  (transS: stored C x k, i.e. the update is Ua*S')
    int k = 0;
    bool transS = false;
};

template <int CP>
__global__ void __launch_bounds__(TSQR_NW * 32, 1) tsqr_reg_kernel(int64_t rows, int C, const double* __restrict__ A, int64_t lda,
                                                                  double* __restrict__ Q, int64_t ldq,
                                                                  double* __restrict__ Rstack, int64_t ldr, TsqrAdd add) {
    using SM = TsqrSmem<CP>;
    constexpr int RPL0 = TSQR_RPL0;
    constexpr int SLD = SM::SROWS + 1;
    constexpr int PROWS = SM::PROWS;
    extern __shared__ __align__(16) double tsm[];
    double* stack = tsm;                          // [CP][SLD]: stacked R factors, later the explicit level-1 Q
    double* stack2 = stack + SM::STACK;           // [CP][SLD]: Householder vectors of the level-1 factorisation
    double* park = stack2 + SM::STACK;            // [NW][CP][PROWS]: per warp: Householder vectors, then the explicit level-0 Q
    double* taus = park + SM::PARK;               // [NW][CP]
    double* bcast = taus + TSQR_NW * CP;          // [NW][2*CP]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* bc = bcast + warp * 2 * CP;
    const int64_t row0 = (int64_t)blockIdx.x * TSQR_BR + warp * PROWS;
    double* mypark = park + (size_t)warp * CP * PROWS;
    {
        double a[RPL0][CP];
        bool ok[RPL0];
#pragma unroll
        for (int q = 0; q < RPL0; ++q) ok[q] = (row0 + lane + 32 * q) < rows;
        const double* col = A + row0 + lane;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
#pragma unroll
            for (int q = 0; q < RPL0; ++q) a[q][c] = (ok[q] && c < C) ? col[32 * q] : 0.0;
            col += lda;
        }
        if (add.U) {
            // stage Sa (k x C) in the (still unused) stack area, then a[q][:] += Ua[row, :] * Sa
            for (int e = threadIdx.x; e < CP * CP; e += blockDim.x) {
                const int kk = e / CP, c = e % CP;
                stack[e] = (kk < add.k && c < C) ? (add.transS ? add.S[c + (int64_t)kk * add.lds] : add.S[kk + (int64_t)c * add.lds]) : 0.0;
            }
            __syncthreads();
#pragma unroll 1
            for (int kk = 0; kk < add.k; ++kk) {
                double u[RPL0];
#pragma unroll
                for (int q = 0; q < RPL0; ++q) u[q] = ok[q] ? add.U[row0 + lane + 32 * q + (int64_t)kk * add.ldu] : 0.0;
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    const double sv = stack[kk * CP + c];
#pragma unroll
                    for (int q = 0; q < RPL0; ++q) a[q][c] = fma(u[q], sv, a[q][c]);
                }
            }
            __syncthreads();   // the stack area is reused for the R factors below
        }
        reg_panel_qr<CP, RPL0>(a, mypark, PROWS, taus + warp * CP, lane, bc);
        // R_w -> rows [CP*warp, CP*warp + CP) of the stack (zeros below the diagonal)
        if (lane < CP) {
#pragma unroll
            for (int c = 0; c < CP; ++c) stack[c * SLD + CP * warp + lane] = (lane <= c) ? mypark[c * PROWS + lane] : 0.0;
        }
        reg_panel_formq<CP, RPL0>(a, mypark, PROWS, taus + warp * CP, lane, bc);
        __syncwarp();
#pragma unroll
        for (int c = 0; c < CP; ++c)
#pragma unroll
            for (int q = 0; q < RPL0; ++q) mypark[c * PROWS + lane + 32 * q] = a[q][c];
    }
    __syncthreads();
    if (warp == 0) tsqr_level1<CP>(stack, stack2, taus, Rstack + (int64_t)blockIdx.x * CP, ldr, lane, bc);
    __syncthreads();
    // Q_w <- Q_w * X_w,  X_w = rows [CP*warp, +CP) of the explicit level-1 Q
#pragma unroll 1
    for (int q = 0; q < RPL0; ++q) {
        double x[CP];
#pragma unroll
        for (int k = 0; k < CP; ++k) x[k] = mypark[k * PROWS + lane + 32 * q];
        const bool okq = (row0 + lane + 32 * q) < rows;
        double* qcol = Q + row0 + lane + 32 * q;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < CP; ++k) s = fma(x[k], stack[c * SLD + CP * warp + k], s);
            if (okq && c < C) *qcol = s;
            qcol += ldq;
        }
    }
}


// ------------------------------------------------------------------------------------------------------
// CTA-wide Householder panel: the 8 warps of a CTA factor ONE 1024 x CP panel together (warp w holds rows 128 w .. 128 w + 127
// in registers, the pivot rows live in warp 0).  Per column the warps' partial dot products meet in shared memory (one
// __syncthreads, double-buffered), every warp forms the same reflector and updates its own rows.  Compared with the
// two-level scheme above (8 independent 128-row panels, then one warp factoring the 8 stacked R's while 7 warps idle, then a
// chain product) a 1024-row block costs ONE panel factorisation + ONE explicit-Q recurrence instead of two of each.
// ------------------------------------------------------------------------------------------------------
template <int CP, int RPL, int NW>
__device__ __forceinline__ void cta_panel_qr(double (&a)[RPL][CP], double* vs, int vld, double* tau_s, int warp, int lane, double* bc,
                                             double* part /*[2][NW][CP]*/, double* prow /*[2][CP]*/) {
    constexpr int SH = (CP == 16) ? 1 : 2;
#pragma unroll 1
    for (int j0 = 0; j0 < CP; j0 += 2) {
        static_for<0, 2>([&](auto PP) {
            constexpr int P = decltype(PP)::value;
            const int j = j0 + P;
            const int par = j & 1;
            double xm[RPL];
#pragma unroll
            for (int q = 0; q < RPL; ++q) xm[q] = ((warp > 0) || (q > 0) || (lane > j)) ? a[q][P] : 0.0;
            double e[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                double s = 0.0;
                if (c + P < CP) {
#pragma unroll
                    for (int q = 0; q < RPL; ++q) s = fma(xm[q], a[q][c + P], s);
                }
                e[c] = s;
            }
            const double h = TReduce<CP, 16>::run(e, lane);
            if ((lane & ((1 << SH) - 1)) == 0) part[(par * NW + warp) * CP + (lane >> SH)] = h;
            if (warp == 0 && lane == j) {
#pragma unroll
                for (int c = P; c < CP; ++c) prow[par * CP + c] = a[0][c];
            }
            __syncthreads();
            if (lane < CP) {   // every warp forms the same totals in the same order
                double tsum = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) tsum += part[(par * NW + w) * CP + lane];
                bc[lane] = tsum;
            }
            __syncwarp();
            const double alpha = prow[par * CP + P];
            const double ss = bc[0];
            double t = 0.0, scale = 0.0, beta = alpha;
            if (ss > 0.0) {
                const double nrm2 = fma(alpha, alpha, ss);
                const double rn = rsqrt(nrm2);
                beta = -copysign(nrm2 * rn, alpha);
                t = fma(fabs(alpha), rn, 1.0);
                scale = 1.0 / (alpha - beta);
            }
            if (warp == 0 && lane == 0) tau_s[j] = t;
            double vm[RPL];
#pragma unroll
            for (int q = 0; q < RPL; ++q) {
                vm[q] = xm[q] * scale;
                a[q][P] = ((warp > 0) || (q > 0) || (lane > j)) ? vm[q] : a[q][P];
            }
            if (warp == 0 && lane == j) { a[0][P] = beta; vm[0] = 1.0; }
#pragma unroll
            for (int q = 0; q < RPL; ++q) vs[j * vld + lane + 32 * q] = a[q][P];
#pragma unroll
            for (int c = P + 1; c < CP; ++c) {
                const double arow = prow[par * CP + c];
                const double ec = bc[c - P];
                const double w = (arow + ec * scale) * t;
#pragma unroll
                for (int q = 0; q < RPL; ++q) a[q][c] = fma(-w, vm[q], a[q][c]);
            }
            __syncwarp();
        });
#pragma unroll
        for (int c = 2; c < CP; ++c)
#pragma unroll
            for (int q = 0; q < RPL; ++q) a[q][c - 2] = a[q][c];
#pragma unroll
        for (int q = 0; q < RPL; ++q) { a[q][CP - 2] = 0.0; a[q][CP - 1] = 0.0; }
    }
    __syncthreads();
}

template <int CP, int RPL, int NW>
__device__ __forceinline__ void cta_panel_formq(double (&a)[RPL][CP], const double* vs, int vld, const double* tau_s, int warp, int lane,
                                                double* bc, double* part) {
    constexpr int SH = (CP == 16) ? 1 : 2;
#pragma unroll
    for (int c = 0; c < CP; ++c)
#pragma unroll
        for (int q = 0; q < RPL; ++q) a[q][c] = 0.0;
#pragma unroll 1
    for (int j0 = CP - 2; j0 >= 0; j0 -= 2) {
#pragma unroll
        for (int c = CP - 1; c >= 2; --c)
#pragma unroll
            for (int q = 0; q < RPL; ++q) a[q][c] = a[q][c - 2];
        static_for<0, 2>([&](auto PP) {
            constexpr int P = 1 - decltype(PP)::value;
            const int j = j0 + P;
            const int par = j & 1;
            const double t = tau_s[j];
            double vm[RPL];
#pragma unroll
            for (int q = 0; q < RPL; ++q) vm[q] = ((warp > 0) || (q > 0) || (lane > j)) ? vs[j * vld + lane + 32 * q] : 0.0;
            double e[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                double s = 0.0;
                if (c + P + 1 < CP) {
#pragma unroll
                    for (int q = 0; q < RPL; ++q) s = fma(vm[q], a[q][c + P + 1], s);
                }
                e[c] = s;
            }
            const double h = TReduce<CP, 16>::run(e, lane);
            if ((lane & ((1 << SH) - 1)) == 0) part[(par * NW + warp) * CP + (lane >> SH)] = h;
            __syncthreads();
            if (lane < CP) {
                double tsum = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) tsum += part[(par * NW + w) * CP + lane];
                bc[lane] = tsum;
            }
            __syncwarp();
            if (warp == 0 && lane == j) vm[0] = 1.0;
#pragma unroll
            for (int c = P + 1; c < CP; ++c) {
                const double w = bc[c - P - 1] * t;
#pragma unroll
                for (int q = 0; q < RPL; ++q) a[q][c] = fma(-w, vm[q], a[q][c]);
            }
#pragma unroll
            for (int q = 0; q < RPL; ++q) a[q][P] = -t * vm[q];
            if (warp == 0 && lane == j) a[0][P] = 1.0 - t;
            __syncwarp();
        });
    }
}

template <int CP>
struct TsqrCtaSmem {
    static constexpr int PROWS = 32 * TSQR_RPL0;
    static constexpr size_t PARK = (size_t)CP * TSQR_NW * PROWS;        // Householder vectors of the 1024 x CP panel
    static constexpr size_t MISC = (size_t)2 * TSQR_NW * CP + 2 * CP + CP + (size_t)TSQR_NW * 2 * CP + CP * CP;
    static constexpr size_t BYTES = (PARK + MISC) * sizeof(double);
};

// one CTA: thin QR of rows [1024 b, 1024 b + 1024) of A -> explicit Q block (in place allowed) + its CP x CP R factor
template <int CP>
__global__ void __launch_bounds__(TSQR_NW * 32, 1) tsqr_cta_kernel(int64_t rows, int C, const double* __restrict__ A, int64_t lda,
                                                                  double* __restrict__ Q, int64_t ldq,
                                                                  double* __restrict__ Rstack, int64_t ldr, TsqrAdd add) {
    using SM = TsqrCtaSmem<CP>;
    constexpr int RPL0 = TSQR_RPL0;
    constexpr int PROWS = SM::PROWS;
    constexpr int VLD = TSQR_NW * PROWS;
    extern __shared__ __align__(16) double tsm[];
    double* park = tsm;                                   // [CP][VLD]
    double* part = park + SM::PARK;                       // [2][NW][CP]
    double* prow = part + 2 * TSQR_NW * CP;               // [2][CP]
    double* taus = prow + 2 * CP;                         // [CP]
    double* bcast = taus + CP;                            // [NW][2*CP]
    double* sadd = bcast + TSQR_NW * 2 * CP;              // [CP][CP] staged Sa of the fused rank-k update
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* bc = bcast + warp * 2 * CP;
    const int64_t row0 = (int64_t)blockIdx.x * TSQR_BR + warp * PROWS;
    double* mypark = park + warp * PROWS;
    double a[RPL0][CP];
    bool ok[RPL0];
#pragma unroll
    for (int q = 0; q < RPL0; ++q) ok[q] = (row0 + lane + 32 * q) < rows;
    {
        const double* col = A + row0 + lane;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
#pragma unroll
            for (int q = 0; q < RPL0; ++q) a[q][c] = (ok[q] && c < C) ? col[32 * q] : 0.0;
            col += lda;
        }
    }
    if (add.U) {
        for (int e = threadIdx.x; e < CP * CP; e += blockDim.x) {
            const int kk = e / CP, c = e % CP;
            sadd[e] = (kk < add.k && c < C) ? (add.transS ? add.S[c + (int64_t)kk * add.lds] : add.S[kk + (int64_t)c * add.lds]) : 0.0;
        }
        __syncthreads();
#pragma unroll 1
        for (int kk = 0; kk < add.k; ++kk) {
            double u[RPL0];
#pragma unroll
            for (int q = 0; q < RPL0; ++q) u[q] = ok[q] ? add.U[row0 + lane + 32 * q + (int64_t)kk * add.ldu] : 0.0;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const double sv = sadd[kk * CP + c];
#pragma unroll
                for (int q = 0; q < RPL0; ++q) a[q][c] = fma(u[q], sv, a[q][c]);
            }
        }
    }
    cta_panel_qr<CP, RPL0, TSQR_NW>(a, mypark, VLD, taus, warp, lane, bc, part, prow);
    if (warp == 0 && lane < CP) {
#pragma unroll
        for (int c = 0; c < CP; ++c) Rstack[(int64_t)blockIdx.x * CP + lane + (int64_t)c * ldr] = (lane <= c) ? park[c * VLD + lane] : 0.0;
    }
    cta_panel_formq<CP, RPL0, TSQR_NW>(a, mypark, VLD, taus, warp, lane, bc, part);
    double* qcol = Q + row0 + lane;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
#pragma unroll
        for (int q = 0; q < RPL0; ++q)
            if (ok[q] && c < C) qcol[32 * q] = a[q][c];
        qcol += ldq;
    }
}

// ------------------------------------------------------------------------------------------------------
// Two-level TSQR in ONE launch (2 <= nb <= 64 row blocks of 1024, i.e. up to 65536 rows per GPU — the headline shape).
// CTAs 0 .. nb-1 factor their panel (cta_panel_qr), publish R, and form their explicit Q block while CTA nb — which only waits
// for the R factors — factors the stacked R's and forms the top Q; the panel CTAs then multiply their register-resident Q block by
// their CP x CP block of the top Q and write the result once.  Against the three launches (level 0, level 1, apply_blocks) the
// sequential chain loses the explicit-Q recurrence of level 0 (it overlaps level 1), the apply launch and one round trip of Q
// through memory.  Arithmetic and summation orders are those of the three kernels (bit-identical results).
// All nb + 1 CTAs must be co-resident while they wait for each other: one CTA per SM, nb + 1 <= 65 of 148 SMs, and nothing that
// runs beside this kernel (auxiliary-stream kernels, the single-GPU exchange-free tail) blocks an SM forever.  The waits are
// bounded (a few seconds) and report a time-out in sync[2] instead of hanging the device.
// sync[0]: running total of published R factors; sync[1]: epoch of the last finished top factorisation.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// thread 0 of the CTA waits until *p has reached `target` (wrap-around compare), then releases the CTA
__device__ __forceinline__ void cta_wait_reached(unsigned int* p, unsigned int target, unsigned int* timeouts) {
    if (threadIdx.x == 0) {
        // plain spin of ONE thread (the others sit at the barrier): a nanosleep between the polls costs tens of microseconds per wait on
        // this part (measured with the LL exchange, profiles/r02/multi_gpu_r02b.txt), far more than the poll traffic it saves
        long long spins = 0;
        while ((int)(ld_acquire_gpu_u32(p) - target) < 0) {
            if (++spins > (1ll << 22)) { atomicAdd(timeouts, 1u); break; }   // a few seconds
        }
    }
    __syncthreads();
}

template <int CP>
__global__ void __launch_bounds__(TSQR_NW * 32, 1) tsqr_fused_kernel(int64_t rows, int C, int nb, const double* __restrict__ A, int64_t lda,
                                                                    double* __restrict__ Q, int64_t ldq, double* __restrict__ Rstack,
                                                                    double* __restrict__ Qtop, double* __restrict__ Rtop, TsqrAdd add,
                                                                    unsigned int* sync, unsigned int arrivals_target, unsigned int epoch) {
    using SM = TsqrCtaSmem<CP>;
    constexpr int RPL0 = TSQR_RPL0;
    constexpr int PROWS = SM::PROWS;
    constexpr int VLD = TSQR_NW * PROWS;
    extern __shared__ __align__(16) double tsm[];
    double* park = tsm;                                   // [CP][VLD]
    double* part = park + SM::PARK;                       // [2][NW][CP]
    double* prow = part + 2 * TSQR_NW * CP;               // [2][CP]
    double* taus = prow + 2 * CP;                         // [CP]
    double* bcast = taus + CP;                            // [NW][2*CP]
    double* sadd = bcast + TSQR_NW * 2 * CP;              // [CP][CP]: staged Sa of the fused rank-k update, later this CTA's block of the top Q
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* bc = bcast + warp * 2 * CP;
    double* mypark = park + warp * PROWS;
    const int64_t ldr = (int64_t)nb * CP;                 // leading dimension of Rstack and Qtop
    double a[RPL0][CP];

    if ((int)blockIdx.x == nb) {
        // ---- top of the tree: wait for the nb R factors, factor the (nb*CP) x CP stack, publish R and the explicit top Q
        cta_wait_reached(sync + 0, arrivals_target, sync + 2);
        const int srows = nb * CP;
        if (srows <= 128) {
            // small stack: one warp, register panel of 128 rows (the arithmetic of tsqr_small_kernel)
            if (warp == 0) {
#pragma unroll
                for (int c = 0; c < CP; ++c)
#pragma unroll
                    for (int q = 0; q < RPL0; ++q) {
                        const int g = lane + 32 * q;
                        a[q][c] = (g < srows) ? __ldcg(Rstack + g + (int64_t)c * ldr) : 0.0;
                    }
                reg_panel_qr<CP, RPL0>(a, park, 32 * RPL0, taus, lane, bc);
                if (lane < CP) {
#pragma unroll
                    for (int c = 0; c < CP; ++c) Rtop[lane + (int64_t)c * CP] = (lane <= c) ? park[c * (32 * RPL0) + lane] : 0.0;
                }
                reg_panel_formq<CP, RPL0>(a, park, 32 * RPL0, taus, lane, bc);
#pragma unroll
                for (int c = 0; c < CP; ++c)
#pragma unroll
                    for (int q = 0; q < RPL0; ++q) {
                        const int g = lane + 32 * q;
                        if (g < srows) Qtop[g + (int64_t)c * ldr] = a[q][c];
                    }
            }
        } else {
            const int r0 = warp * PROWS + lane;
#pragma unroll
            for (int c = 0; c < CP; ++c)
#pragma unroll
                for (int q = 0; q < RPL0; ++q) {
                    const int g = r0 + 32 * q;
                    a[q][c] = (g < srows) ? __ldcg(Rstack + g + (int64_t)c * ldr) : 0.0;
                }
            cta_panel_qr<CP, RPL0, TSQR_NW>(a, mypark, VLD, taus, warp, lane, bc, part, prow);
            if (warp == 0 && lane < CP) {
#pragma unroll
                for (int c = 0; c < CP; ++c) Rtop[lane + (int64_t)c * CP] = (lane <= c) ? park[c * VLD + lane] : 0.0;
            }
            cta_panel_formq<CP, RPL0, TSQR_NW>(a, mypark, VLD, taus, warp, lane, bc, part);
#pragma unroll
            for (int c = 0; c < CP; ++c)
#pragma unroll
                for (int q = 0; q < RPL0; ++q) {
                    const int g = r0 + 32 * q;
                    if (g < srows) Qtop[g + (int64_t)c * ldr] = a[q][c];
                }
        }
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); st_release_gpu_u32(sync + 1, epoch); }
        return;
    }

    // ---- panel CTA: rows [1024 b, 1024 b + 1024)
    const int64_t row0 = (int64_t)blockIdx.x * TSQR_BR + warp * PROWS;
    bool ok[RPL0];
#pragma unroll
    for (int q = 0; q < RPL0; ++q) ok[q] = (row0 + lane + 32 * q) < rows;
    {
        const double* col = A + row0 + lane;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
#pragma unroll
            for (int q = 0; q < RPL0; ++q) a[q][c] = (ok[q] && c < C) ? col[32 * q] : 0.0;
            col += lda;
        }
    }
    if (add.U) {
        for (int e = threadIdx.x; e < CP * CP; e += blockDim.x) {
            const int kk = e / CP, c = e % CP;
            sadd[e] = (kk < add.k && c < C) ? (add.transS ? add.S[c + (int64_t)kk * add.lds] : add.S[kk + (int64_t)c * add.lds]) : 0.0;
        }
        __syncthreads();
#pragma unroll 1
        for (int kk = 0; kk < add.k; ++kk) {
            double u[RPL0];
#pragma unroll
            for (int q = 0; q < RPL0; ++q) u[q] = ok[q] ? add.U[row0 + lane + 32 * q + (int64_t)kk * add.ldu] : 0.0;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const double sv = sadd[kk * CP + c];
#pragma unroll
                for (int q = 0; q < RPL0; ++q) a[q][c] = fma(u[q], sv, a[q][c]);
            }
        }
    }
    cta_panel_qr<CP, RPL0, TSQR_NW>(a, mypark, VLD, taus, warp, lane, bc, part, prow);
    if (warp == 0 && lane < CP) {
#pragma unroll
        for (int c = 0; c < CP; ++c) Rstack[(int64_t)blockIdx.x * CP + lane + (int64_t)c * ldr] = (lane <= c) ? park[c * VLD + lane] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(sync + 0, 1u); }     // this panel's R is published
    cta_panel_formq<CP, RPL0, TSQR_NW>(a, mypark, VLD, taus, warp, lane, bc, part);   // overlaps the top factorisation
    // park the explicit Q block in shared memory (the Householder vectors there are dead; every thread rereads only its own
    // entries): the register panel is free while this CTA waits and during the product below
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CP; ++c)
#pragma unroll
        for (int q = 0; q < RPL0; ++q) mypark[c * VLD + lane + 32 * q] = a[q][c];
    cta_wait_reached(sync + 1, epoch, sync + 2);
    // this CTA's CP x CP block X_b of the top Q:  Q_block <- Q_block * X_b  (the arithmetic of apply_blocks_kernel)
    for (int e = threadIdx.x; e < CP * CP; e += blockDim.x) {
        const int k = e % CP, c = e / CP;
        sadd[k * CP + c] = (k < C && c < C) ? __ldcg(Qtop + (int64_t)blockIdx.x * CP + k + (int64_t)c * ldr) : 0.0;
    }
    __syncthreads();
#pragma unroll 1
    for (int q = 0; q < RPL0; ++q) {
        double x[CP];
#pragma unroll
        for (int k = 0; k < CP; ++k) x[k] = (k < C) ? mypark[k * VLD + lane + 32 * q] : 0.0;
        double* qcol = Q + row0 + lane + 32 * q;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < CP; ++k) s = fma(x[k], sadd[k * CP + c], s);
            if (ok[q] && c < C) *qcol = s;
            qcol += ldq;
        }
    }
}

// rows <= 128: one warp factors the whole matrix (used for the top of the tree and for all-gathered R stacks)
template <int CP>
__global__ void __launch_bounds__(32, 1) tsqr_small_kernel(int rows, int C, const double* __restrict__ A, int64_t lda,
                                                          double* __restrict__ Q, int64_t ldq, double* __restrict__ Rout, int64_t ldr) {
    constexpr int RPL = 4;
    __shared__ double vs[CP][32 * RPL];
    __shared__ double taus[CP];
    __shared__ double bc[2 * CP];
    const int lane = threadIdx.x;
    double a[RPL][CP];
#pragma unroll
    for (int c = 0; c < CP; ++c)
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            const int g = lane + 32 * q;
            a[q][c] = (g < rows && c < C) ? A[g + (int64_t)c * lda] : 0.0;
        }
    reg_panel_qr<CP, RPL>(a, &vs[0][0], 32 * RPL, taus, lane, bc);
    if (lane < CP) {
#pragma unroll
        for (int c = 0; c < CP; ++c) Rout[lane + (int64_t)c * ldr] = (lane <= c) ? vs[c][lane] : 0.0;
    }
    reg_panel_formq<CP, RPL>(a, &vs[0][0], 32 * RPL, taus, lane, bc);
#pragma unroll
    for (int c = 0; c < CP; ++c)
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            const int g = lane + 32 * q;
            if (g < rows && c < C) Q[g + (int64_t)c * ldq] = a[q][c];
        }
}

// Cross-rank top of the TSQR tree in ONE single-warp launch (P2P transport): post this rank's R (CP x CP, ld CP) to the exchange
// region, raise the flag in every peer, wait for the peers, read their R factors straight into the register panel
// (rows g*CP.. of the stacked G*CP x CP matrix, over NVLink), factor it, and write the replicated R plus THIS rank's CP x CP
// block of the explicit Q.  Every rank factors the same stack in the same order => bit-identical R on all ranks.
template <int CP, bool LL>
__global__ void __launch_bounds__(32, 1) tsqr_xrank_kernel(P2PView v, LLView lv, const double* __restrict__ Rloc, double* __restrict__ Rout,
                                                          double* __restrict__ Xout) {
    constexpr int RPL = 4;   // G*CP <= 128 rows
    __shared__ double vs[CP][32 * RPL];
    __shared__ double taus[CP];
    __shared__ double bc[2 * CP];
    const int lane = threadIdx.x;
    const int nranks = LL ? lv.nranks : v.nranks, rank = LL ? lv.rank : v.rank;
    if (LL) {
        for (int e = lane; e < CP * CP; e += 32) ll_push(lv, e, Rloc[e]);   // flag-in-data: R goes straight into every rank's buffer
        ll_flush(lv);
    } else {
        for (int e = lane; e < CP * CP; e += 32) v.data_local[e] = Rloc[e];
        __syncwarp();
        __threadfence_system();
        if (lane < nranks) st_release_sys(v.flags_peer[lane] + (size_t)rank * P2P_FLAG_STRIDE, v.seq);
        if (lane < nranks) {
            const unsigned long long* f = v.flags_local + (size_t)lane * P2P_FLAG_STRIDE;
            while (ld_acquire_sys(f) < v.seq) { }
        }
        __syncwarp();
    }
    const int rows = nranks * CP;
    double a[RPL][CP];
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        const int gr = lane + 32 * q;
        const int g = gr / CP, i = gr % CP;
        if (LL) {
            // all CP loads of a row are issued before the first flag is looked at (one L2 round trip per attempt, not CP)
            bool ok;
            do {
                ok = true;
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    if (gr >= rows) a[q][c] = 0.0;
                    else ok &= ll_try_load(lv.local + (size_t)g * lv.cap + i + c * CP, lv.seq, a[q][c], lv.atomic_poll);
                }
                if (!ok && lv.backoff > 0) __nanosleep(lv.backoff);
            } while (!ok);
        } else {
#pragma unroll
            for (int c = 0; c < CP; ++c) a[q][c] = (gr < rows) ? ld_relaxed_sys(v.data_peer[g] + i + c * CP) : 0.0;
        }
    }
    reg_panel_qr<CP, RPL>(a, &vs[0][0], 32 * RPL, taus, lane, bc);
    if (lane < CP) {
#pragma unroll
        for (int c = 0; c < CP; ++c) Rout[lane + c * CP] = (lane <= c) ? vs[c][lane] : 0.0;
    }
    reg_panel_formq<CP, RPL>(a, &vs[0][0], 32 * RPL, taus, lane, bc);
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        const int gr = lane + 32 * q;
        if (gr >= rank * CP && gr < (rank + 1) * CP) {
#pragma unroll
            for (int c = 0; c < CP; ++c) Xout[(gr - rank * CP) + c * CP] = a[q][c];
        }
    }
}

// Q[rows of block b, :C] <- Q[block b, :C] * X_b[:C,:C],  X_b = Xstack[b*xstride .. , :] (ldx); block = block_rows rows.
template <int CP>
__global__ void __launch_bounds__(128) apply_blocks_kernel(int64_t rows, int C, double* __restrict__ Q, int64_t ldq, int64_t block_rows,
                                                          const double* __restrict__ Xstack, int64_t ldx, int64_t xstride) {
    __shared__ double Xs[CP][CP + 1];
    const int64_t base = (int64_t)blockIdx.x * 128;
    const int64_t b = base / block_rows;  // 128 | block_rows
    for (int e = threadIdx.x; e < CP * CP; e += 128) {
        int k = e % CP, c = e / CP;
        Xs[k][c] = (k < C && c < C) ? Xstack[b * xstride + k + (int64_t)c * ldx] : 0.0;
    }
    __syncthreads();
    const int64_t i = base + threadIdx.x;
    if (i >= rows) return;
    double x[CP];
#pragma unroll
    for (int k = 0; k < CP; ++k) x[k] = (k < C) ? Q[i + (int64_t)k * ldq] : 0.0;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < CP; ++k) s = fma(x[k], Xs[k][c], s);
        if (c < C) Q[i + (int64_t)c * ldq] = s;
    }
}

inline int tsqr_cp(int C) { return C <= 8 ? 8 : 16; }

// doubles of scratch for a rows x C (C <= 16) TSQR incl. recursion and the optional NCCL gather stage
inline int64_t tsqr_ws_size(int64_t rows, int C, int nranks) {
    const int CP = tsqr_cp(C);
    int64_t total = 0, r = rows;
    while (true) {
        int64_t nb = cdiv(r, TSQR_BR);
        total += 2 * nb * CP * CP;  // Rstack + Qtop of this level
        if (nb == 1) break;
        r = nb * CP;
    }
    total += (int64_t)(nranks + 1) * CP * CP * 4;
    return total + 1024;
}

inline void tsqr_level(Ctx& cx, int CP, int64_t rows, int C, const double* A, int64_t lda, double* Q, int64_t ldq, double* Rstack, int64_t ldr,
                       const TsqrAdd& add = TsqrAdd()) {
    if (rows <= 128) {
        DLRA_REQUIRE(add.U == nullptr, "fused rank-k update is only available on the multi-warp TSQR level");
        if (CP == 8) tsqr_small_kernel<8><<<1, 32, 0, cx.stream>>>((int)rows, C, A, lda, Q, ldq, Rstack, ldr);
        else tsqr_small_kernel<16><<<1, 32, 0, cx.stream>>>((int)rows, C, A, lda, Q, ldq, Rstack, ldr);
        cx.launches++;
        DLRA_CUDA(cudaGetLastError());
        return;
    }
    const unsigned nb = (unsigned)cdiv(rows, TSQR_BR);
    static const bool legacy = getenv("DLRA_TSQR_LEGACY") != nullptr;
    if (!legacy) {
        static unsigned long long attr_devs_c = 0;
        if (first_use_on_this_device(attr_devs_c)) {
            DLRA_CUDA(cudaFuncSetAttribute(tsqr_cta_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsqrCtaSmem<8>::BYTES));
            DLRA_CUDA(cudaFuncSetAttribute(tsqr_cta_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsqrCtaSmem<16>::BYTES));
        }
        if (CP == 8) tsqr_cta_kernel<8><<<nb, TSQR_NW * 32, TsqrCtaSmem<8>::BYTES, cx.stream>>>(rows, C, A, lda, Q, ldq, Rstack, ldr, add);
        else tsqr_cta_kernel<16><<<nb, TSQR_NW * 32, TsqrCtaSmem<16>::BYTES, cx.stream>>>(rows, C, A, lda, Q, ldq, Rstack, ldr, add);
        cx.launches++;
        DLRA_CUDA(cudaGetLastError());
        return;
    }
    static unsigned long long attr_devs = 0;
    if (first_use_on_this_device(attr_devs)) {
        DLRA_CUDA(cudaFuncSetAttribute(tsqr_reg_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsqrSmem<8>::BYTES));
        DLRA_CUDA(cudaFuncSetAttribute(tsqr_reg_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsqrSmem<16>::BYTES));
    }
    if (CP == 8) tsqr_reg_kernel<8><<<nb, TSQR_NW * 32, TsqrSmem<8>::BYTES, cx.stream>>>(rows, C, A, lda, Q, ldq, Rstack, ldr, add);
    else tsqr_reg_kernel<16><<<nb, TSQR_NW * 32, TsqrSmem<16>::BYTES, cx.stream>>>(rows, C, A, lda, Q, ldq, Rstack, ldr, add);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

inline void apply_blocks(Ctx& cx, int CP, int64_t rows, int C, double* Q, int64_t ldq, int64_t block_rows, const double* X, int64_t ldx, int64_t xstride) {
    unsigned grid = (unsigned)cdiv(rows, 128);
    if (CP == 8) apply_blocks_kernel<8><<<grid, 128, 0, cx.stream>>>(rows, C, Q, ldq, block_rows, X, ldx, xstride);
    else apply_blocks_kernel<16><<<grid, 128, 0, cx.stream>>>(rows, C, Q, ldq, block_rows, X, ldx, xstride);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

// Local (single GPU) TSQR of A (rows x C, C <= 16): Q (rows x C, may alias A).  Returns a pointer (inside ws) to the
// CP x CP padded R factor (ld = CP).  ws must hold tsqr_ws_size doubles.
inline double* tsqr_local(Ctx& cx, int64_t rows, int C, const double* A, int64_t lda, double* Q, int64_t ldq, double* ws,
                          const TsqrAdd& add = TsqrAdd()) {
    const int CP = tsqr_cp(C);
    const int64_t nb = cdiv(rows, TSQR_BR);
    double* Rstack = ws;                    // (nb*CP) x CP, ld = nb*CP
    double* Qtop = ws + nb * CP * CP;       // same shape
    double* rest = Qtop + nb * CP * CP;
    static const bool fused = !(getenv("DLRA_TSQR_FUSED") && atoi(getenv("DLRA_TSQR_FUSED")) == 0) && getenv("DLRA_TSQR_LEGACY") == nullptr;
    // measured (profiles/r02/tsqr_fused_ab.txt): 64 blocks 87.5 -> 77-81 us.  With 4 blocks (m = 4096) the single launch is slower than the
    // three small ones in isolation (68 vs 57 us), but the m-side chain is not the critical one in the BUG step and three small launches beside
    // the n-side kernel slow THAT one down (81.1 vs 76.7 us): BUG 1.114 -> 1.107 ms, KSL 1.507 -> 1.511 ms.  Default: every tree of 2..64 blocks.
    static const int fused_min_nb = getenv("DLRA_TSQR_FUSED_MIN") ? atoi(getenv("DLRA_TSQR_FUSED_MIN")) : 2;
    if (fused && nb >= fused_min_nb && nb >= 2 && nb <= 64 && nb + 1 <= cx.num_sms / 2 && cx.sync != nullptr) {
        static unsigned long long attr_devs_f = 0;
        if (first_use_on_this_device(attr_devs_f)) {
            DLRA_CUDA(cudaFuncSetAttribute(tsqr_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsqrCtaSmem<8>::BYTES));
            DLRA_CUDA(cudaFuncSetAttribute(tsqr_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsqrCtaSmem<16>::BYTES));
        }
        cx.sync_arrivals += (unsigned int)nb;
        cx.sync_epoch += 1;
        double* Rtop = rest;                // CP x CP, ld = CP (where the recursion below would put it)
        if (CP == 8) tsqr_fused_kernel<8><<<(unsigned)nb + 1, TSQR_NW * 32, TsqrCtaSmem<8>::BYTES, cx.stream>>>(rows, C, (int)nb, A, lda, Q, ldq, Rstack, Qtop, Rtop, add, cx.sync, cx.sync_arrivals, cx.sync_epoch);
        else tsqr_fused_kernel<16><<<(unsigned)nb + 1, TSQR_NW * 32, TsqrCtaSmem<16>::BYTES, cx.stream>>>(rows, C, (int)nb, A, lda, Q, ldq, Rstack, Qtop, Rtop, add, cx.sync, cx.sync_arrivals, cx.sync_epoch);
        cx.launches++;
        DLRA_CUDA(cudaGetLastError());
        return Rtop;
    }
    tsqr_level(cx, CP, rows, C, A, lda, Q, ldq, Rstack, nb * CP, add);
    if (nb == 1) return Rstack;
    double* Rtop = tsqr_local(cx, nb * CP, CP, Rstack, nb * CP, Qtop, nb * CP, rest);
    apply_blocks(cx, CP, rows, C, Q, ldq, TSQR_BR, Qtop, nb * CP, CP);
    return Rtop;
}

// Distributed TSQR: rows are this rank's shard.  R (C x C, ldr) optional output (replicated on all ranks).
inline void tsqr(Ctx& cx, Comm& comm, int64_t rows, int C, const double* A, int64_t lda, double* Q, int64_t ldq, double* R, int64_t ldr,
                 double* ws, const TsqrAdd& add = TsqrAdd()) {
    DLRA_REQUIRE(C >= 1 && C <= TSQR_MAXC, "tsqr panel width must be 1..16");
    const int CP = tsqr_cp(C);
    double* tail = ws + tsqr_ws_size(rows, C, comm.nranks) - 1024 - (int64_t)(comm.nranks + 1) * CP * CP * 4;
    double* Rloc = tsqr_local(cx, rows, C, A, lda, Q, ldq, ws, add);
    const double* Rfin = Rloc;
    if (comm.nranks > 1) cx.mark("tsqr_local");
    if (comm.nranks > 1 && comm.p2p) {
        DLRA_REQUIRE(comm.nranks * CP <= 128, "too many ranks for the single-warp cross-rank R reduction");
        DLRA_REQUIRE((size_t)CP * CP * 8 <= comm.xdata_bytes, "P2P exchange region too small for an R factor");
        double* Rg = tail;                 // CP x CP replicated R
        double* X = tail + CP * CP;        // this rank's block of the stacked Q
        if (comm.ll_fits((int64_t)CP * CP)) {
            LLView lv = comm.next_ll(0);
            if (CP == 8) tsqr_xrank_kernel<8, true><<<1, 32, 0, cx.stream>>>(P2PView{}, lv, Rloc, Rg, X);
            else tsqr_xrank_kernel<16, true><<<1, 32, 0, cx.stream>>>(P2PView{}, lv, Rloc, Rg, X);
        } else {
            P2PView v = comm.next_view(0);
            if (CP == 8) tsqr_xrank_kernel<8, false><<<1, 32, 0, cx.stream>>>(v, LLView{}, Rloc, Rg, X);
            else tsqr_xrank_kernel<16, false><<<1, 32, 0, cx.stream>>>(v, LLView{}, Rloc, Rg, X);
        }
        cx.launches++;
        DLRA_CUDA(cudaGetLastError());
        cx.mark("tsqr_xrank");
        apply_blocks(cx, CP, rows, C, Q, ldq, (int64_t)1 << 62, X, CP, 0);
        cx.mark("tsqr_xapply");
        Rfin = Rg;
    } else if (comm.nranks > 1) {
        const int G = comm.nranks;
        double* gathered = tail;                                // G blocks of CP x CP (each ld = CP)
        double* stacked = gathered + (int64_t)G * CP * CP;      // (G*CP) x CP, ld = G*CP
        double* Qg = stacked + (int64_t)G * CP * CP;            // (G*CP) x CP
        double* ws2 = Qg + (int64_t)G * CP * CP;
        comm.allgather(Rloc, gathered, (int64_t)CP * CP, cx);
        for (int g = 0; g < G; ++g) copy_mat(cx, CP, CP, gathered + (int64_t)g * CP * CP, CP, false, stacked + (int64_t)g * CP, (int64_t)G * CP);
        DLRA_REQUIRE((int64_t)G * CP <= TSQR_BR, "too many ranks for the single-CTA R reduction");
        double* Rg = tsqr_local(cx, (int64_t)G * CP, CP, stacked, (int64_t)G * CP, Qg, (int64_t)G * CP, ws2);
        apply_blocks(cx, CP, rows, C, Q, ldq, (int64_t)1 << 62, Qg + (int64_t)comm.rank * CP, (int64_t)G * CP, 0);
        Rfin = Rg;
    }
    if (R) copy_mat(cx, C, C, Rfin, CP, false, R, ldr);
}

// Wide thin QR (any C): block classical Gram-Schmidt with re-orthogonalisation around <=16-column TSQR panels.
// Per block: project, TSQR, project again, TSQR again (Barlow-Smoktunowicz BCGS2) — the intermediate factorisation makes
// the second projection act on a perfectly conditioned block, so ||I - Q'Q|| stays O(eps) even when the block is almost
// contained in the span of the previous ones (the rank-adaptive [K U0] basis always is).
// A is overwritten; Q may alias A.  R (C x C, ldr) optional.  gws: gemm_tn scratch, Wtmp: thin_qr_wtmp(C) doubles, tws: tsqr scratch.
inline int64_t thin_qr_wtmp(int C) { return (int64_t)2 * C * TSQR_MAXC + 4 * TSQR_MAXC * TSQR_MAXC; }

// ortho_cols: number of leading columns of A that are already orthonormal (panels entirely inside them are taken as they are).
inline void thin_qr(Ctx& cx, Comm& comm, int64_t rows, int C, double* A, int64_t lda, double* Q, int64_t ldq, double* R, int64_t ldr,
                    double* tws, double* gws, double* Wtmp, int ortho_cols = 0) {
    if (C <= TSQR_MAXC) {
        tsqr(cx, comm, rows, C, A, lda, Q, ldq, R, ldr, tws);
        return;
    }
    double* W1 = Wtmp;
    double* W2 = W1 + (int64_t)C * TSQR_MAXC;
    double* R1 = W2 + (int64_t)C * TSQR_MAXC;
    double* R2 = R1 + TSQR_MAXC * TSQR_MAXC;
    constexpr int LR = TSQR_MAXC;
    if (R) fill_mat(cx, C, C, R, ldr, 0.0, 0.0);
    for (int c0 = 0; c0 < C; c0 += TSQR_MAXC) {
        const int cb = std::min(TSQR_MAXC, C - c0);
        double* Ap = A + (int64_t)c0 * lda;
        double* Qp = Q + (int64_t)c0 * ldq;
        if (c0 + cb <= ortho_cols) {   // already orthonormal (and orthogonal to the earlier panels, which lie in the same block)
            if (Qp != Ap) copy_mat(cx, rows, cb, Ap, lda, false, Qp, ldq);
            if (R) fill_mat(cx, cb, cb, R + c0 + (int64_t)c0 * ldr, ldr, 0.0, 1.0);
            continue;
        }
        if (c0 == 0) {
            tsqr(cx, comm, rows, cb, Ap, lda, Qp, ldq, R, ldr, tws);
            continue;
        }
        // first pass: W1 = Qprev' * Ap ; Ap -= Qprev * W1 ; Ap = Qhat * R1
        gemm_tn(cx, rows, c0, cb, Q, ldq, nullptr, 0, Ap, lda, W1, c0, 1.0, 0.0, gws);
        comm.allreduce_sum(W1, (int64_t)c0 * cb, cx);
        gemm_nn(cx, rows, c0, cb, Q, ldq, nullptr, 0, W1, c0, false, Ap, lda, -1.0, 1.0);
        tsqr(cx, comm, rows, cb, Ap, lda, Qp, ldq, R1, LR, tws);
        // second pass on the orthonormal block: W2 = Qprev' * Qhat ; Qhat -= Qprev * W2 ; Qhat = Qp * R2
        gemm_tn(cx, rows, c0, cb, Q, ldq, nullptr, 0, Qp, ldq, W2, c0, 1.0, 0.0, gws);
        comm.allreduce_sum(W2, (int64_t)c0 * cb, cx);
        gemm_nn(cx, rows, c0, cb, Q, ldq, nullptr, 0, W2, c0, false, Qp, ldq, -1.0, 1.0);
        tsqr(cx, comm, rows, cb, Qp, ldq, Qp, ldq, R ? R2 : nullptr, LR, tws);
        if (R) {
            // A_p = Qprev*(W1 + W2*R1) + Qp*(R2*R1)
            small_gemm(cx, c0, cb, cb, W2, c0, false, R1, LR, false, W1, c0, 1.0, 1.0);
            copy_mat(cx, c0, cb, W1, c0, false, R + (int64_t)c0 * ldr, ldr);
            small_gemm(cx, cb, cb, cb, R2, LR, false, R1, LR, false, R + c0 + (int64_t)c0 * ldr, (int)ldr, 1.0, 0.0);
        }
    }
}

}  // namespace dlra
