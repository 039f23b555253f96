// Communication-avoiding tall-skinny Householder QR (TSQR) with explicit thin Q.
// Replaces `qr!(US)` + `Matrix(QRK.Q)` (LAPACK dgeqrt + dgemqrt) at projector_splitting.jl:137-138,149-151,
// 174-175,186-188; unconventional.jl:141-142,149-152; rank_adaptive_unconventional.jl:202-205,213-216.
//
// Structure (one kernel per tree level, all levels Householder => unconditionally stable, rank-deficient
// panels included):
//   tsqr_cta_kernel : each CTA takes BR = NW0*64 rows; every warp factors a 64 x CP panel held in shared
//                     memory with warp-shuffle reductions (panel_qr), forms its explicit Q in place
//                     (panel_formq, dorg2r-style), the R factors are stacked and reduced by an in-CTA tree,
//                     then the explicit Q's are chained top-down so the CTA writes an orthonormal BR x C
//                     block plus ONE CP x CP R factor.
//   recursion       : the stacked R factors (nb*CP x CP) are factored by the same kernel until one CTA
//                     suffices; with row sharding the per-GPU R's are all-gathered (NCCL) and every rank
//                     redundantly factors the G*CP x CP stack.
//   apply_blocks    : Q_block <- Q_block * X_block with the CP x CP blocks of the upper level's Q.
// Columns wider than 32 are handled by block classical Gram-Schmidt with re-orthogonalisation (BCGS2) around
// 32-column TSQR panels.
#pragma once
#include "common.cuh"
#include "comm.cuh"
#include "small_ops.cuh"

namespace dlra {

constexpr int TSQR_PLD = 65;  // panel leading dimension (64 rows + 1 pad): conflict-light for row- and column-parallel access

template <int CP>
struct TsqrCfg {
    static constexpr int NW0 = (CP == 32) ? 4 : 8;   // level-0 warps (panels) per CTA
    static constexpr int BR = NW0 * 64;              // rows per CTA
    static constexpr int ARITY = 64 / CP;            // R factors stacked per upper-level panel
    static constexpr int cnt(int level) { int c = NW0; for (int l = 0; l < level; ++l) c = (c * CP + 63) / 64; return c; }
    static constexpr int levels() { int l = 1, c = NW0; while (c > 1) { c = (c * CP + 63) / 64; ++l; } return l; }
    static constexpr int offset(int level) { int o = 0; for (int l = 0; l < level; ++l) o += cnt(l); return o; }
    static constexpr int total() { return offset(levels()); }
    static constexpr int prows(int level) { return level == 0 ? 64 : ((cnt(level - 1) * CP >= 64) ? 64 : ((cnt(level - 1) * CP + 31) / 32) * 32); }
    static constexpr size_t smem_bytes() { return (size_t)total() * CP * TSQR_PLD * 8 + (size_t)total() * CP * 8 + (size_t)NW0 * CP * 8; }
};

// Householder QR of a prows x CP panel P[c*PLD + row] by one warp. R ends in the upper triangle, the
// Householder vectors (unit diagonal implicit) below it, tau[j] as LAPACK dlarfg (tau = 0 for a zero column).
template <int CP>
__device__ void panel_qr(double* P, int prows, double* tau, double* wbuf, int lane) {
    constexpr int PLD = TSQR_PLD;
    for (int j = 0; j < CP; ++j) {
        double* colj = P + j * PLD;
        double ss = 0.0;
        for (int row = lane; row < prows; row += 32)
            if (row > j) { double x = colj[row]; ss = fma(x, x, ss); }
        ss = warp_sum(ss);
        const double alpha = colj[j];
        double t = 0.0, scale = 0.0, beta = alpha;
        if (ss > 0.0) {
            beta = -copysign(sqrt(fma(alpha, alpha, ss)), alpha);
            t = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        __syncwarp();
        for (int row = lane; row < prows; row += 32)
            if (row > j) colj[row] *= scale;
        if (lane == 0) { colj[j] = beta; tau[j] = t; }
        __syncwarp();
        if (t != 0.0) {
            // w_c = tau * v' * P[:,c]   (lane <-> column)
            for (int c = j + 1 + lane; c < CP; c += 32) {
                const double* colc = P + c * PLD;
                double w0 = colc[j], w1 = 0.0, w2 = 0.0, w3 = 0.0;
                int row = j + 1;
                for (; row + 3 < prows; row += 4) {
                    w0 = fma(colj[row], colc[row], w0);
                    w1 = fma(colj[row + 1], colc[row + 1], w1);
                    w2 = fma(colj[row + 2], colc[row + 2], w2);
                    w3 = fma(colj[row + 3], colc[row + 3], w3);
                }
                for (; row < prows; ++row) w0 = fma(colj[row], colc[row], w0);
                wbuf[c] = ((w0 + w1) + (w2 + w3)) * t;
            }
            __syncwarp();
            // P[:,c] -= v * w_c       (lane <-> row)
            for (int c = j + 1; c < CP; ++c) {
                const double w = wbuf[c];
                double* colc = P + c * PLD;
                for (int row = lane; row < prows; row += 32) {
                    if (row > j) colc[row] = fma(-w, colj[row], colc[row]);
                    else if (row == j) colc[row] -= w;
                }
            }
            __syncwarp();
        }
    }
}

// In-place explicit thin Q (prows x CP) from the Householder vectors left by panel_qr (LAPACK dorg2r).
// The upper triangle (R) must have been copied out before.
template <int CP>
__device__ void panel_formq(double* P, int prows, const double* tau, double* wbuf, int lane) {
    constexpr int PLD = TSQR_PLD;
    for (int j = CP - 1; j >= 0; --j) {
        double* colj = P + j * PLD;
        const double t = tau[j];
        // rows < j of columns > j are already zero; row j of columns > j currently holds R -> must read as 0
        if (t != 0.0 && j + 1 < CP) {
            for (int c = j + 1 + lane; c < CP; c += 32) {
                const double* colc = P + c * PLD;
                double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;  // Q[j][c] == 0 before H_j is applied
                int row = j + 1;
                for (; row + 3 < prows; row += 4) {
                    w0 = fma(colj[row], colc[row], w0);
                    w1 = fma(colj[row + 1], colc[row + 1], w1);
                    w2 = fma(colj[row + 2], colc[row + 2], w2);
                    w3 = fma(colj[row + 3], colc[row + 3], w3);
                }
                for (; row < prows; ++row) w0 = fma(colj[row], colc[row], w0);
                wbuf[c] = ((w0 + w1) + (w2 + w3)) * t;
            }
            __syncwarp();
            for (int c = j + 1; c < CP; ++c) {
                const double w = wbuf[c];
                double* colc = P + c * PLD;
                for (int row = lane; row < prows; row += 32) {
                    if (row > j) colc[row] = fma(-w, colj[row], colc[row]);
                    else if (row == j) colc[row] = -w;
                }
            }
        } else {
            for (int c = j + 1 + lane; c < CP; c += 32) P[c * PLD + j] = 0.0;
        }
        __syncwarp();
        for (int row = lane; row < prows; row += 32) {
            if (row > j) colj[row] = -t * colj[row];
            else if (row == j) colj[row] = 1.0 - t;
            else colj[row] = 0.0;
        }
        __syncwarp();
    }
}

template <int CP>
__global__ void __launch_bounds__(TsqrCfg<CP>::NW0 * 32) tsqr_cta_kernel(int64_t rows, int C, const double* __restrict__ A, int64_t lda,
                                                                       double* __restrict__ Q, int64_t ldq,
                                                                       double* __restrict__ Rstack, int64_t ldr) {
    using Cfg = TsqrCfg<CP>;
    constexpr int PLD = TSQR_PLD;
    constexpr int NL = Cfg::levels();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* panels = reinterpret_cast<double*>(smem_raw);
    double* taus = panels + (size_t)Cfg::total() * CP * PLD;
    double* wbufs = taus + (size_t)Cfg::total() * CP;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * Cfg::BR;
    double* wbuf = wbufs + warp * CP;

    // ---- level 0 load (zero padded rows / columns)
    {
        double* P = panels + (size_t)warp * CP * PLD;
        for (int c = 0; c < CP; ++c)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int row = lane + 32 * q;
                const int64_t g = row0 + warp * 64 + row;
                P[c * PLD + row] = (g < rows && c < C) ? A[g + (int64_t)c * lda] : 0.0;
            }
    }
    __syncwarp();
    // ---- up-sweep
#pragma unroll
    for (int lev = 0; lev < NL; ++lev) {
        const int cnt = Cfg::cnt(lev);
        if (warp < cnt) {
            double* P = panels + (size_t)(Cfg::offset(lev) + warp) * CP * PLD;
            double* tau = taus + (size_t)(Cfg::offset(lev) + warp) * CP;
            const int prows = Cfg::prows(lev);
            panel_qr<CP>(P, prows, tau, wbuf, lane);
            // copy R (upper triangle, zeros below) to the parent panel or to the global stack
            if (lev + 1 < NL) {
                double* Pp = panels + (size_t)(Cfg::offset(lev + 1) + warp / Cfg::ARITY) * CP * PLD;
                const int roff = (warp % Cfg::ARITY) * CP;
                for (int e = lane; e < CP * CP; e += 32) {
                    int i = e % CP, c = e / CP;
                    Pp[c * PLD + roff + i] = (i <= c) ? P[c * PLD + i] : 0.0;
                }
            } else {
                for (int e = lane; e < CP * CP; e += 32) {
                    int i = e % CP, c = e / CP;
                    Rstack[(int64_t)blockIdx.x * CP + i + (int64_t)c * ldr] = (i <= c) ? P[c * PLD + i] : 0.0;
                }
            }
            __syncwarp();
            panel_formq<CP>(P, prows, tau, wbuf, lane);
        }
        __syncthreads();
    }
    // ---- down-sweep: Q_panel <- Q_panel * X, X = CP x CP row block of the parent's (already final) panel
#pragma unroll
    for (int lev = NL - 2; lev >= 0; --lev) {
        const int cnt = Cfg::cnt(lev);
        if (warp < cnt) {
            double* P = panels + (size_t)(Cfg::offset(lev) + warp) * CP * PLD;
            const double* X = panels + (size_t)(Cfg::offset(lev + 1) + warp / Cfg::ARITY) * CP * PLD + (warp % Cfg::ARITY) * CP;
            const int prows = Cfg::prows(lev);
            for (int row = lane; row < prows; row += 32) {
                double x[CP];
#pragma unroll
                for (int k = 0; k < CP; ++k) x[k] = P[k * PLD + row];
                for (int c = 0; c < CP; ++c) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < CP; ++k) s = fma(x[k], X[c * PLD + k], s);
                    P[c * PLD + row] = s;
                }
            }
        }
        __syncthreads();
    }
    // ---- store the explicit Q block
    {
        const double* P = panels + (size_t)warp * CP * PLD;
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int row = lane + 32 * q;
                const int64_t g = row0 + warp * 64 + row;
                if (g < rows) Q[g + (int64_t)c * ldq] = P[c * PLD + row];
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
    for (int c = 0; c < C; ++c) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < CP; ++k) s = fma(x[k], Xs[k][c], s);
        Q[i + (int64_t)c * ldq] = s;
    }
}

struct TsqrWorkspace {
    double* buf = nullptr;  // device doubles
    int64_t size = 0;
};

inline int tsqr_cp(int C) { return C <= 8 ? 8 : (C <= 16 ? 16 : 32); }
inline int64_t tsqr_br(int CP) { return CP == 32 ? TsqrCfg<32>::BR : TsqrCfg<16>::BR; }

// doubles of scratch for a rows x C (C <= 32) TSQR incl. recursion and the optional NCCL gather stage
inline int64_t tsqr_ws_size(int64_t rows, int C, int nranks) {
    const int CP = tsqr_cp(C);
    const int64_t BR = tsqr_br(CP);
    int64_t total = 0, r = rows;
    while (true) {
        int64_t nb = cdiv(r, BR);
        total += 2 * nb * CP * CP;  // Rstack + Qtop of this level
        if (nb == 1) break;
        r = nb * CP;
    }
    total += (int64_t)(nranks + 1) * CP * CP * 4;
    return total + 1024;
}

template <int CP>
inline void tsqr_launch(Ctx& cx, int64_t rows, int C, const double* A, int64_t lda, double* Q, int64_t ldq, double* Rstack, int64_t ldr) {
    using Cfg = TsqrCfg<CP>;
    static bool attr_set = false;
    if (!attr_set) {
        DLRA_CUDA(cudaFuncSetAttribute(tsqr_cta_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes()));
        attr_set = true;
    }
    const int64_t nb = cdiv(rows, Cfg::BR);
    tsqr_cta_kernel<CP><<<(unsigned)nb, Cfg::NW0 * 32, Cfg::smem_bytes(), cx.stream>>>(rows, C, A, lda, Q, ldq, Rstack, ldr);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

inline void tsqr_level(Ctx& cx, int CP, int64_t rows, int C, const double* A, int64_t lda, double* Q, int64_t ldq, double* Rstack, int64_t ldr) {
    if (CP == 8) tsqr_launch<8>(cx, rows, C, A, lda, Q, ldq, Rstack, ldr);
    else if (CP == 16) tsqr_launch<16>(cx, rows, C, A, lda, Q, ldq, Rstack, ldr);
    else tsqr_launch<32>(cx, rows, C, A, lda, Q, ldq, Rstack, ldr);
}

inline void apply_blocks(Ctx& cx, int CP, int64_t rows, int C, double* Q, int64_t ldq, int64_t block_rows, const double* X, int64_t ldx, int64_t xstride) {
    unsigned grid = (unsigned)cdiv(rows, 128);
    if (CP == 8) apply_blocks_kernel<8><<<grid, 128, 0, cx.stream>>>(rows, C, Q, ldq, block_rows, X, ldx, xstride);
    else if (CP == 16) apply_blocks_kernel<16><<<grid, 128, 0, cx.stream>>>(rows, C, Q, ldq, block_rows, X, ldx, xstride);
    else apply_blocks_kernel<32><<<grid, 128, 0, cx.stream>>>(rows, C, Q, ldq, block_rows, X, ldx, xstride);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

// Local (single GPU) TSQR of A (rows x C, C <= 32): Q (rows x C, may alias A) and Rp = CP x CP padded R (ld = CP).
// ws must hold tsqr_ws_size doubles.  Returns a pointer (inside ws) to the CP x CP R factor.
inline double* tsqr_local(Ctx& cx, int64_t rows, int C, const double* A, int64_t lda, double* Q, int64_t ldq, double* ws) {
    const int CP = tsqr_cp(C);
    const int64_t BR = tsqr_br(CP);
    const int64_t nb = cdiv(rows, BR);
    double* Rstack = ws;                    // (nb*CP) x CP, ld = nb*CP
    double* Qtop = ws + nb * CP * CP;       // same shape
    double* rest = Qtop + nb * CP * CP;
    tsqr_level(cx, CP, rows, C, A, lda, Q, ldq, Rstack, nb * CP);
    if (nb == 1) return Rstack;
    double* Rtop = tsqr_local(cx, nb * CP, CP, Rstack, nb * CP, Qtop, nb * CP, rest);
    apply_blocks(cx, CP, rows, C, Q, ldq, BR, Qtop, nb * CP, CP);
    return Rtop;
}

// Distributed TSQR: rows are this rank's shard.  R (C x C, ldr) optional output (replicated on all ranks).
inline void tsqr(Ctx& cx, Comm& comm, int64_t rows, int C, const double* A, int64_t lda, double* Q, int64_t ldq, double* R, int64_t ldr,
                 double* ws) {
    DLRA_REQUIRE(C >= 1 && C <= 32, "tsqr panel width must be 1..32");
    const int CP = tsqr_cp(C);
    double* tail = ws + tsqr_ws_size(rows, C, comm.nranks) - 1024 - (int64_t)(comm.nranks + 1) * CP * CP * 4;
    double* Rloc = tsqr_local(cx, rows, C, A, lda, Q, ldq, ws);
    const double* Rfin = Rloc;
    if (comm.nranks > 1) {
        const int G = comm.nranks;
        double* gathered = tail;                       // G blocks of CP x CP (each ld = CP)
        double* stacked = gathered + (int64_t)G * CP * CP;      // (G*CP) x CP, ld = G*CP
        double* Qg = stacked + (int64_t)G * CP * CP;            // (G*CP) x CP
        double* ws2 = Qg + (int64_t)G * CP * CP;                // CP*CP*? small recursion scratch (G*CP <= BR assumed)
        comm.allgather(Rloc, gathered, (int64_t)CP * CP, cx.stream);
        for (int g = 0; g < G; ++g) copy_mat(cx, CP, CP, gathered + (int64_t)g * CP * CP, CP, false, stacked + (int64_t)g * CP, (int64_t)G * CP);
        DLRA_REQUIRE((int64_t)G * CP <= tsqr_br(CP), "too many ranks for the single-CTA R reduction");
        double* Rg = tsqr_local(cx, (int64_t)G * CP, CP, stacked, (int64_t)G * CP, Qg, (int64_t)G * CP, ws2);
        apply_blocks(cx, CP, rows, C, Q, ldq, (int64_t)1 << 62, Qg + (int64_t)comm.rank * CP, (int64_t)G * CP, 0);
        Rfin = Rg;
    }
    if (R) copy_mat(cx, C, C, Rfin, CP, false, R, ldr);
}

// Wide thin QR (any C): BCGS2 around <=32-column TSQR panels.  A is overwritten; Q may alias A.
// R (C x C, ldr) optional.  gws: gemm_tn scratch, Wtmp: C x 32 scratch (device), tws: tsqr scratch.
inline void thin_qr(Ctx& cx, Comm& comm, int64_t rows, int C, double* A, int64_t lda, double* Q, int64_t ldq, double* R, int64_t ldr,
                    double* tws, double* gws, double* Wtmp) {
    if (C <= 32) {
        tsqr(cx, comm, rows, C, A, lda, Q, ldq, R, ldr, tws);
        return;
    }
    if (R) fill_mat(cx, C, C, R, ldr, 0.0, 0.0);
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int cb = std::min(32, C - c0);
        double* Ap = A + (int64_t)c0 * lda;
        if (c0 > 0) {
            for (int pass = 0; pass < 2; ++pass) {
                // W = Q[:, :c0]' * Ap  (c0 x cb), all-reduced over row shards
                gemm_tn(cx, rows, c0, cb, Q, ldq, nullptr, 0, Ap, lda, Wtmp, c0, 1.0, 0.0, gws);
                comm.allreduce_sum(Wtmp, (int64_t)c0 * cb, cx.stream);
                gemm_nn(cx, rows, c0, cb, Q, ldq, nullptr, 0, Wtmp, c0, false, Ap, lda, -1.0, 1.0);
                if (R) copy_mat(cx, c0, cb, Wtmp, c0, false, R + (int64_t)c0 * ldr, ldr, 1.0, pass == 0 ? 0.0 : 1.0);
            }
        }
        tsqr(cx, comm, rows, cb, Ap, lda, Q + (int64_t)c0 * ldq, ldq, R ? R + c0 + (int64_t)c0 * ldr : nullptr, ldr, tws);
    }
}

}  // namespace dlra
