// Generic fp64 helpers: tall-skinny GEMMs (any shape / alignment), tiny dense ops, norms.
// These carry (a) every O((n+m) r^2) product of a step: U*S, V*S', M = U1'U0, N = V1'V0, M*S*N', Uhat*P
// (reference: mul! at projector_splitting.jl:133,145,170,182; unconventional.jl:137,142,145,150,154;
// rank_adaptive_unconventional.jl:198,206,209,217,219,225,227,231) and (b) the generic fallback of the
// K/L/S contractions `u .+= sign*left'*Δy*right` (data_integrator.jl:13-16) for shapes the TMA/DMMA
// fast path (pass_tma.cuh) does not take.  All matrices column-major.
#pragma once
#include "common.cuh"

namespace dlra {

// ------------------------------------------------------------------------------------------------
// C[n x q] = beta*C + alpha * (A - Aprev)[n x p] * op(B)[p x q]        (tall A, small op(B); one thread per row)
// transB: op(B)[k][c] = B[c + k*ldb]  (B stored q x p)
// ------------------------------------------------------------------------------------------------
template <int QC>
__global__ void __launch_bounds__(128) gemm_nn_kernel(int64_t n, int p, int q, const double* __restrict__ A, int64_t lda,
                                                       const double* __restrict__ Aprev, int64_t ldap,
                                                       const double* __restrict__ B, int64_t ldb, int transB,
                                                       double* __restrict__ C, int64_t ldc, double alpha, double beta) {
    constexpr int KT = QC >= 64 ? 32 : 64;
    __shared__ double Bs[KT][QC];
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int c0 = blockIdx.y * QC;
    double acc[QC];
#pragma unroll
    for (int c = 0; c < QC; ++c) acc[c] = 0.0;
    for (int k0 = 0; k0 < p; k0 += KT) {
        __syncthreads();
        for (int e = threadIdx.x; e < KT * QC; e += 128) {
            int k = e / QC, c = e % QC;
            double v = 0.0;
            if (k0 + k < p && c0 + c < q) v = transB ? B[(c0 + c) + (int64_t)(k0 + k) * ldb] : B[(k0 + k) + (int64_t)(c0 + c) * ldb];
            Bs[k][c] = v;
        }
        __syncthreads();
        if (i < n) {
            const int kmax = min(KT, p - k0);
            for (int k = 0; k < kmax; ++k) {
                double a = A[i + (int64_t)(k0 + k) * lda];
                if (Aprev) a -= Aprev[i + (int64_t)(k0 + k) * ldap];
#pragma unroll
                for (int c = 0; c < QC; ++c) acc[c] = fma(a, Bs[k][c], acc[c]);
            }
        }
    }
    if (i < n) {
#pragma unroll
        for (int c = 0; c < QC; ++c)
            if (c0 + c < q) {
                double* dst = C + i + (int64_t)(c0 + c) * ldc;
                *dst = (beta == 0.0 ? 0.0 : beta * (*dst)) + alpha * acc[c];
            }
    }
}

inline void gemm_nn(Ctx& cx, int64_t n, int p, int q, const double* A, int64_t lda, const double* Aprev, int64_t ldap,
                    const double* B, int64_t ldb, bool transB, double* C, int64_t ldc, double alpha, double beta) {
    if (n <= 0 || q <= 0) return;
    if (!Aprev && cx.tall_gemm && cx.tall_gemm(cx.tall_eng, n, p, q, A, lda, B, ldb, transB, C, ldc, alpha, beta)) return;
    if (q > 16 && n >= 8192) {
        dim3 grid((unsigned)cdiv(n, 128), (unsigned)cdiv(q, 32));
        gemm_nn_kernel<32><<<grid, 128, 0, cx.stream>>>(n, p, q, A, lda, Aprev, ldap, B, ldb, transB ? 1 : 0, C, ldc, alpha, beta);
    } else {
        dim3 grid((unsigned)cdiv(n, 128), (unsigned)cdiv(q, 16));
        gemm_nn_kernel<16><<<grid, 128, 0, cx.stream>>>(n, p, q, A, lda, Aprev, ldap, B, ldb, transB ? 1 : 0, C, ldc, alpha, beta);
    }
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Cpart[chunk][p x q] = (A - Aprev)[rows of chunk, p]' * B[rows of chunk, q]     (long reduction over rows)
// followed by a fixed-order reduction over chunks (deterministic; no atomics).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_tn_part_kernel(int64_t n, int p, int q, int64_t chunk_rows,
                                                           const double* __restrict__ A, int64_t lda,
                                                           const double* __restrict__ Aprev, int64_t ldap,
                                                           const double* __restrict__ B, int64_t ldb,
                                                           double* __restrict__ Cpart) {
    constexpr int RT = 64, PB = 32, QB = 32;
    __shared__ double As[PB][RT + 1];
    __shared__ double Bs[QB][RT + 1];
    const int a0 = blockIdx.y * PB, b0 = blockIdx.z * QB;
    const int64_t r0 = (int64_t)blockIdx.x * chunk_rows;
    const int64_t r1 = min(n, r0 + chunk_rows);
    const int ta = threadIdx.x % 16, tb = threadIdx.x / 16;
    double acc00 = 0, acc01 = 0, acc10 = 0, acc11 = 0;
    const int lr = threadIdx.x % RT, lc = threadIdx.x / RT;  // 4 column groups of 8
    for (int64_t rr = r0; rr < r1; rr += RT) {
        __syncthreads();
        const int64_t i = rr + lr;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            int c = lc * 8 + cc;
            double va = 0.0, vb = 0.0;
            if (i < r1) {
                if (a0 + c < p) {
                    va = A[i + (int64_t)(a0 + c) * lda];
                    if (Aprev) va -= Aprev[i + (int64_t)(a0 + c) * ldap];
                }
                if (b0 + c < q) vb = B[i + (int64_t)(b0 + c) * ldb];
            }
            As[c][lr] = va;
            Bs[c][lr] = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < RT; ++k) {
            double x0 = As[ta][k], x1 = As[ta + 16][k], y0 = Bs[tb][k], y1 = Bs[tb + 16][k];
            acc00 = fma(x0, y0, acc00);
            acc01 = fma(x0, y1, acc01);
            acc10 = fma(x1, y0, acc10);
            acc11 = fma(x1, y1, acc11);
        }
    }
    double* out = Cpart + (int64_t)blockIdx.x * p * q;
    if (a0 + ta < p && b0 + tb < q) out[(a0 + ta) + (int64_t)(b0 + tb) * p] = acc00;
    if (a0 + ta < p && b0 + tb + 16 < q) out[(a0 + ta) + (int64_t)(b0 + tb + 16) * p] = acc01;
    if (a0 + ta + 16 < p && b0 + tb < q) out[(a0 + ta + 16) + (int64_t)(b0 + tb) * p] = acc10;
    if (a0 + ta + 16 < p && b0 + tb + 16 < q) out[(a0 + ta + 16) + (int64_t)(b0 + tb + 16) * p] = acc11;
}

// C[p x q] (ldc) = beta*C + alpha * sum_chunks part[chunk]   (part: dense p x q per chunk, leading dim ldp)
__global__ void reduce_parts_kernel(int p, int q, int nchunks, const double* __restrict__ part, int64_t ldp, int64_t chunk_stride,
                                    double* __restrict__ C, int64_t ldc, double alpha, double beta) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)p * q) return;
    int a = (int)(e % p), b = (int)(e / p);
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += part[(int64_t)c * chunk_stride + a + (int64_t)b * ldp];
    double* dst = C + a + (int64_t)b * ldc;
    *dst = (beta == 0.0 ? 0.0 : beta * (*dst)) + alpha * s;
}

inline void reduce_parts(Ctx& cx, int p, int q, int nchunks, const double* part, int64_t ldp, int64_t chunk_stride, double* C,
                         int64_t ldc, double alpha, double beta) {
    int64_t tot = (int64_t)p * q;
    if (tot <= 0) return;
    reduce_parts_kernel<<<(unsigned)cdiv(tot, 256), 256, 0, cx.stream>>>(p, q, nchunks, part, ldp, chunk_stride, C, ldc, alpha, beta);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Gram-type product for small outputs on the fp64 tensor pipe, ONE launch:
//   C[16-block (y), 16-block (z)] = beta*C + alpha * A[:, y-block]' * B[:, z-block]      (rows = long reduction)
// Each warp strides over 4-row slabs (DMMA k = 4; fragments are 32-byte runs of the column-major operands), the 8 warps'
// tiles are summed in shared memory, every CTA stores one partial and the LAST CTA to finish (ticket counter) adds the
// partials in fixed order — deterministic without a second launch.
// ------------------------------------------------------------------------------------------------
// Sum of base[b*stride], b = 0 .. count-1, in index order, with 16 loads in flight per thread: the last-CTA reductions below are
// chains of L2 round trips (one thread owns one output element), so their time is count / (loads in flight) x the L2 latency.
__device__ __forceinline__ double ordered_sum_strided(const double* base, int count, int64_t stride) {
    double s = 0.0;
    int b = 0;
    for (; b + 16 <= count; b += 16) {
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldcg(base + (int64_t)(b + u) * stride);
#pragma unroll
        for (int u = 0; u < 16; ++u) s += v[u];
    }
#pragma unroll 4
    for (; b < count; ++b) s += __ldcg(base + (int64_t)b * stride);
    return s;
}

constexpr int GRAM_ROWS_PER_CTA = 256;
__global__ void __launch_bounds__(256) gram_dmma_kernel(int64_t n, int p, int q, const double* __restrict__ A, int64_t lda,
                                                        const double* __restrict__ B, int64_t ldb, double* __restrict__ C, int64_t ldc,
                                                        double alpha, double beta, double* __restrict__ part, unsigned int* __restrict__ counters) {
    __shared__ double red[8][256];
    __shared__ bool is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, k = lane & 3;
    const int a0 = blockIdx.y * 16, b0 = blockIdx.z * 16;
    const int64_t r0 = (int64_t)blockIdx.x * GRAM_ROWS_PER_CTA, r1 = min(n, r0 + GRAM_ROWS_PER_CTA);
    double acc[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
    const bool a_ok[2] = {a0 + g < p, a0 + 8 + g < p};
    const bool b_ok[2] = {b0 + g < q, b0 + 8 + g < q};
    const double* Ap[2] = {A + (int64_t)(a0 + g) * lda, A + (int64_t)(a0 + 8 + g) * lda};
    const double* Bp[2] = {B + (int64_t)(b0 + g) * ldb, B + (int64_t)(b0 + 8 + g) * ldb};
#pragma unroll 4
    for (int64_t row = r0 + 4 * warp; row < r1; row += 32) {
        const int64_t i = row + k;
        const bool ok = i < r1;
        double af[2], bf[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            af[h] = (ok && a_ok[h]) ? Ap[h][i] : 0.0;
            bf[h] = (ok && b_ok[h]) ? Bp[h][i] : 0.0;
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], af[mb], bf[nb]);
    }
    // C fragment: rows (mb*8 + g), cols (nb*8 + 2k + e)
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) red[warp][(mb * 8 + g) + 16 * (nb * 8 + 2 * k + e)] = acc[mb][nb][e];
    __syncthreads();
    const int nblk = gridDim.x;
    const int slot = blockIdx.y * gridDim.z + blockIdx.z;
    double* mypart = part + ((int64_t)slot * nblk + blockIdx.x) * 256;
    {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        mypart[threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(&counters[slot], 1u);
        is_last = (ticket == (unsigned int)nblk - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const double s = ordered_sum_strided(part + (int64_t)slot * nblk * 256 + threadIdx.x, nblk, 256);   // block order
    const int i = threadIdx.x % 16, j = threadIdx.x / 16;
    if (a0 + i < p && b0 + j < q) {
        double* dst = C + (a0 + i) + (int64_t)(b0 + j) * ldc;
        *dst = (beta == 0.0 ? 0.0 : beta * (*dst)) + alpha * s;
    }
    if (threadIdx.x == 0) counters[slot] = 0;   // self-reset for the next launch on this stream
}

// The tail of the pipelined BUG step in ONE launch (single GPU):  Rm = A'·B  (A = U1, B = ΔA·V1: the core increment, r <= 16) with the
// arithmetic of gram_dmma_kernel, and — in the last CTA to finish — the core update  Out = M·S0·N' + Rm  with the arithmetic of
// core_update_kernel (unconventional.jl:154-155).  Saves a launch and the round trip of Rm through global memory on the critical path
// between the streaming pass and the next step's QR.  Out may alias S0.
constexpr int GRAM_CORE_GROUP = 16;
// scratch (doubles) and the largest row count the group counters (255 of the 256 per stream) allow
inline int64_t gram_core_ws(int64_t n) { const int64_t nb = cdiv(n, GRAM_ROWS_PER_CTA); return (nb + cdiv(nb, GRAM_CORE_GROUP)) * 256; }
inline bool gram_core_ok(int64_t n) { return cdiv(cdiv(n, GRAM_ROWS_PER_CTA), GRAM_CORE_GROUP) <= 255; }
__global__ void __launch_bounds__(256) gram_core_kernel(int64_t n, int r, const double* __restrict__ A, int64_t lda,
                                                        const double* __restrict__ B, int64_t ldb, double* __restrict__ Rm,
                                                        double* __restrict__ part, unsigned int* __restrict__ counters,
                                                        const double* __restrict__ M, const double* S0,
                                                        const double* __restrict__ Nn, double* Out, int ld) {
    constexpr int GC = GRAM_CORE_GROUP;
    __shared__ double red[8][256];
    __shared__ bool is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, k = lane & 3;
    const int64_t r0 = (int64_t)blockIdx.x * GRAM_ROWS_PER_CTA, r1 = min(n, r0 + GRAM_ROWS_PER_CTA);
    double acc[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
    const bool c_ok[2] = {g < r, 8 + g < r};
    const double* Ap[2] = {A + (int64_t)g * lda, A + (int64_t)(8 + g) * lda};
    const double* Bp[2] = {B + (int64_t)g * ldb, B + (int64_t)(8 + g) * ldb};
#pragma unroll 4
    for (int64_t row = r0 + 4 * warp; row < r1; row += 32) {
        const int64_t i = row + k;
        const bool ok = i < r1;
        double af[2], bf[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            af[h] = (ok && c_ok[h]) ? Ap[h][i] : 0.0;
            bf[h] = (ok && c_ok[h]) ? Bp[h][i] : 0.0;
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], af[mb], bf[nb]);
    }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) red[warp][(mb * 8 + g) + 16 * (nb * 8 + 2 * k + e)] = acc[mb][nb][e];
    __syncthreads();
    const int nblk = gridDim.x;
    {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        part[(int64_t)blockIdx.x * 256 + threadIdx.x] = s;
    }
    // Two-level fixed-order reduction (a single CTA summing all nblk partials is a chain of nblk/16 L2 round trips): the last CTA of
    // every group of GC consecutive CTAs sums its group (one batch of loads), the last group to finish sums the group sums.
    // counters[0]: finished groups; counters[1 + g]: finished CTAs of group g (all self-resetting).  The summation tree depends on
    // the CTA indices only, never on the order of arrival.
    const int ngrp = (nblk + GC - 1) / GC;
    const int grp = blockIdx.x / GC;
    const int gsize = min(GC, nblk - grp * GC);
    double* gpart = part + (int64_t)nblk * 256;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(&counters[1 + grp], 1u);
        is_last = (ticket == (unsigned int)gsize - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    gpart[(int64_t)grp * 256 + threadIdx.x] = ordered_sum_strided(part + (int64_t)grp * GC * 256 + threadIdx.x, gsize, 256);
    if (threadIdx.x == 0) counters[1 + grp] = 0;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(&counters[0], 1u);
        is_last = (ticket == (unsigned int)ngrp - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // operands of the core update: staged while the partial sums are in flight
    double (*sM)[17] = reinterpret_cast<double (*)[17]>(&red[0][0]);        // red is dead: 4 x 16 x 17 doubles fit in its first rows
    double (*sS)[17] = sM + 16;
    double (*sN)[17] = sS + 16;
    double (*sT)[17] = sN + 16;
    const int i = threadIdx.x % 16, j = threadIdx.x / 16;
    const bool in = i < r && j < r;
    sM[i][j] = in ? M[i + (int64_t)j * ld] : 0.0;
    sS[i][j] = in ? S0[i + (int64_t)j * ld] : 0.0;
    sN[i][j] = in ? Nn[i + (int64_t)j * ld] : 0.0;
    double s = ordered_sum_strided(gpart + threadIdx.x, ngrp, 256);   // fixed group order
    s = 0.0 + 1.0 * s;
    if (in) Rm[i + (int64_t)j * ld] = s;
    __syncthreads();
    {   // T = M*S0 (thread <-> T[i][j])
        double t = 0.0;
        for (int l = 0; l < r; ++l) t = fma(sM[i][l], sS[l][j], t);
        sT[i][j] = t;
    }
    __syncthreads();
    {
        double t = 0.0;
        for (int l = 0; l < r; ++l) t = fma(sT[i][l], sN[j][l], t);
        if (in) Out[i + (int64_t)j * ld] = t + s;
    }
    if (threadIdx.x == 0) counters[0] = 0;
}
inline void gram_core(Ctx& cx, int64_t n, int r, const double* A, int64_t lda, const double* B, int64_t ldb, double* Rm, double* ws,
                      const double* M, const double* S0, const double* Nn, double* Out, int ld) {
    gram_core_kernel<<<(unsigned)cdiv(n, GRAM_ROWS_PER_CTA), 256, 0, cx.stream>>>(n, r, A, lda, B, ldb, Rm, ws, cx.counters, M, S0, Nn, Out, ld);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}



// ------------------------------------------------------------------------------------------------
// Wide Gram-type product (p or q > 16, long n) on the fp64 tensor pipe, ONE launch, A and B streamed ONCE per 64 x 64 output
// block:  a persistent CTA per SM owns a contiguous range of 32-row tiles, stages [32 rows x 64 cols] of A and of B through a
// 4-deep cp.async ring (column stride 36 doubles: both DMMA fragment patterns are bank-conflict free) and its 8 warps
// (2 x 4, 32 x 16 outputs each) run DMMA.8x8x4 over them.  Per-CTA partials are reduced in fixed order by the last CTA
// (ticket counter) — deterministic, no atomics on data.  Needs 16-byte aligned column starts (even lda/ldb).
// ------------------------------------------------------------------------------------------------
constexpr int GT_ROWS = 32, GT_LD = 36, GT_BLK = 64, GT_STAGES = 4;
template <int QB> struct GramTileCfg {   // QB = 64: 64 x 64 output block per CTA; QB = 16: 64 x 16 (BCGS2 projections: q <= 16)
    static constexpr int TILE_A = GT_BLK * GT_LD, TILE_B = QB * GT_LD;
    static constexpr int SMEM_BYTES = GT_STAGES * (TILE_A + TILE_B) * (int)sizeof(double);
    static constexpr int MB = QB == 64 ? 4 : 1;   // 8-row blocks of the output per warp
};

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = smem_u32(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int QB>
__global__ void __launch_bounds__(256, 1) gram_tile_kernel(int64_t n, int p, int q, const double* __restrict__ A, int64_t lda,
                                                           const double* __restrict__ B, int64_t ldb, double* __restrict__ C, int64_t ldc,
                                                           double alpha, double beta, double* __restrict__ part,
                                                           unsigned int* __restrict__ counters, int64_t tiles_per_cta) {
    using CF = GramTileCfg<QB>;
    constexpr int MB = CF::MB;
    extern __shared__ __align__(16) double gts[];
    __shared__ bool is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, k = lane & 3;
    // warp tile: QB = 64: rows [32 wm, +32) x cols [16 wn, +16) (2 x 4 warps); QB = 16: rows [8 warp, +8) x all 16 cols
    const int row_w = QB == 64 ? (warp >> 2) * 32 : warp * 8;
    const int col_w = QB == 64 ? (warp & 3) * 16 : 0;
    const int a0 = blockIdx.y * GT_BLK, b0 = blockIdx.z * QB;
    const int64_t ntiles = (n + GT_ROWS - 1) / GT_ROWS;
    const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta, t1 = min(ntiles, t0 + tiles_per_cta);
    double acc[MB][2][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) { acc[mb][nb][0] = 0.0; acc[mb][nb][1] = 0.0; }

    // loader: an operand tile is (64 | QB) columns x 16 chunks of 16 bytes
    auto issue = [&](int64_t t, int stage) {
        double* As = gts + (size_t)stage * (CF::TILE_A + CF::TILE_B);
        double* Bs = As + CF::TILE_A;
        const int64_t r0 = t * GT_ROWS;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int e = threadIdx.x + it * 256;
            const int col = e >> 4, ch = e & 15;
            const int64_t row = r0 + 2 * ch;
            const bool aok = row < n && (a0 + col) < p;
            const double* ga = A + (aok ? row + (int64_t)(a0 + col) * lda : 0);
            cp_async16_zfill(As + col * GT_LD + 2 * ch, ga, aok);
            if (col < QB) {
                const bool bok = row < n && (b0 + col) < q;
                const double* gb = B + (bok ? row + (int64_t)(b0 + col) * ldb : 0);
                cp_async16_zfill(Bs + col * GT_LD + 2 * ch, gb, bok);
            }
        }
    };
    const int64_t my = t1 > t0 ? t1 - t0 : 0;
#pragma unroll
    for (int s = 0; s < GT_STAGES - 1; ++s) {
        if (s < my) issue(t0 + s, s);
        cp_async_commit();
    }
    for (int64_t i = 0; i < my; ++i) {
        cp_async_wait<GT_STAGES - 2>();
        __syncthreads();                       // tile i landed for everybody; everybody is done with tile i-1's slot
        if (i + GT_STAGES - 1 < my) issue(t0 + i + GT_STAGES - 1, (int)((i + GT_STAGES - 1) % GT_STAGES));
        cp_async_commit();
        const double* As = gts + (size_t)(i % GT_STAGES) * (CF::TILE_A + CF::TILE_B);
        const double* Bs = As + CF::TILE_A;
        const double* ap = As + (row_w + g) * GT_LD + k;
        const double* bp = Bs + (col_w + g) * GT_LD + k;
#pragma unroll
        for (int ks = 0; ks < GT_ROWS / 4; ++ks) {
            double af[MB], bf[2];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) af[mb] = ap[mb * 8 * GT_LD + ks * 4];
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) bf[nb] = bp[nb * 8 * GT_LD + ks * 4];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], af[mb], bf[nb]);
        }
    }
    cp_async_wait<0>();
    const int nblk = gridDim.x;
    const int slot = blockIdx.y * gridDim.z + blockIdx.z;
    constexpr int PART = GT_BLK * QB;
    double* mypart = part + ((int64_t)slot * nblk + blockIdx.x) * PART;
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                mypart[(row_w + mb * 8 + g) + GT_BLK * (col_w + nb * 8 + 2 * k + e)] = acc[mb][nb][e];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(&counters[slot], 1u);
        is_last = (ticket == (unsigned int)nblk - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const double* base = part + (int64_t)slot * nblk * PART;
    for (int e = threadIdx.x; e < PART; e += 256) {
        const int i = e % GT_BLK, j = e / GT_BLK;
        if (a0 + i >= p || b0 + j >= q) continue;
        const double s = ordered_sum_strided(base + e, nblk, PART);
        double* dst = C + (a0 + i) + (int64_t)(b0 + j) * ldc;
        *dst = (beta == 0.0 ? 0.0 : beta * (*dst)) + alpha * s;
    }
    if (threadIdx.x == 0) counters[slot] = 0;
}

inline bool gram_tile_ok(const Ctx& cx, int64_t n, int p, int q, const double* A, int64_t lda, const double* B, int64_t ldb) {
    if (cx.counters == nullptr || n < 8192 || (p <= 16 && q <= 16)) return false;
    if (cdiv(p, GT_BLK) * cdiv(q, GT_BLK) > 256) return false;
    return (n % 2) == 0 && (lda % 2) == 0 && (ldb % 2) == 0 && (((uintptr_t)A) & 15) == 0 && (((uintptr_t)B) & 15) == 0;
}
inline int gram_tile_ctas(const Ctx& cx, int64_t n) { return (int)std::min<int64_t>(cx.num_sms, cdiv(n, GT_ROWS)); }

// workspace (doubles) needed by gemm_tn for the partial sums
inline int64_t gemm_tn_chunks(const Ctx& cx, int64_t n, int p, int q) {
    int64_t blocks = cdiv(p, 32) * cdiv(q, 32);
    int64_t want = cdiv((int64_t)4 * cx.num_sms, blocks);
    int64_t chunk_rows = round_up(cdiv(n, want), 64);
    if (chunk_rows < 256) chunk_rows = 256;
    return cdiv(n, chunk_rows);
}
inline bool gram_fast_ok(const Ctx& cx, int p, int q) { return cx.counters != nullptr && cdiv(p, 16) * cdiv(q, 16) <= 256; }
inline int64_t gemm_tn_ws(const Ctx& cx, int64_t n, int p, int q) {
    const int64_t generic = gemm_tn_chunks(cx, n, p, q) * p * q;
    const int64_t fast = cdiv(p, 16) * cdiv(q, 16) * cdiv(n, GRAM_ROWS_PER_CTA) * 256;
    const int64_t tiled = cdiv(p, GT_BLK) * (int64_t)gram_tile_ctas(cx, n) * GT_BLK * (q <= 16 ? 16 : cdiv(q, GT_BLK) * GT_BLK);
    return std::max(std::max(generic, fast), tiled);
}

// C[p x q] = beta*C + alpha * (A - Aprev)' * B
inline void gemm_tn(Ctx& cx, int64_t n, int p, int q, const double* A, int64_t lda, const double* Aprev, int64_t ldap,
                    const double* B, int64_t ldb, double* C, int64_t ldc, double alpha, double beta, double* ws) {
    if (p <= 0 || q <= 0) return;
    if (!Aprev && gram_tile_ok(cx, n, p, q, A, lda, B, ldb)) {
        static unsigned long long attr_devs = 0;
        if (first_use_on_this_device(attr_devs)) {
            DLRA_CUDA(cudaFuncSetAttribute(gram_tile_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GramTileCfg<64>::SMEM_BYTES));
            DLRA_CUDA(cudaFuncSetAttribute(gram_tile_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GramTileCfg<16>::SMEM_BYTES));
        }
        const int G = gram_tile_ctas(cx, n);
        const int64_t tiles_per_cta = cdiv(cdiv(n, GT_ROWS), G);
        if (q <= 16) {   // narrow right-hand side (BCGS2 projections): a 64 x 16 block per CTA, a quarter of the DMMA work of the square block
            dim3 grid((unsigned)G, (unsigned)cdiv(p, GT_BLK), 1);
            gram_tile_kernel<16><<<grid, 256, GramTileCfg<16>::SMEM_BYTES, cx.stream>>>(n, p, q, A, lda, B, ldb, C, ldc, alpha, beta, ws, cx.counters, tiles_per_cta);
        } else {
            dim3 grid((unsigned)G, (unsigned)cdiv(p, GT_BLK), (unsigned)cdiv(q, GT_BLK));
            gram_tile_kernel<64><<<grid, 256, GramTileCfg<64>::SMEM_BYTES, cx.stream>>>(n, p, q, A, lda, B, ldb, C, ldc, alpha, beta, ws, cx.counters, tiles_per_cta);
        }
        cx.launches++;
        DLRA_CUDA(cudaGetLastError());
        return;
    }
    if (!Aprev && gram_fast_ok(cx, p, q)) {
        dim3 grid((unsigned)cdiv(n, GRAM_ROWS_PER_CTA), (unsigned)cdiv(p, 16), (unsigned)cdiv(q, 16));
        gram_dmma_kernel<<<grid, 256, 0, cx.stream>>>(n, p, q, A, lda, B, ldb, C, ldc, alpha, beta, ws, cx.counters);
        cx.launches++;
        DLRA_CUDA(cudaGetLastError());
        return;
    }
    int64_t nch = gemm_tn_chunks(cx, n, p, q);
    int64_t chunk_rows = round_up(cdiv(n, nch), 64);
    nch = cdiv(n, chunk_rows);
    if (nch < 1) nch = 1;
    dim3 grid((unsigned)nch, (unsigned)cdiv(p, 32), (unsigned)cdiv(q, 32));
    gemm_tn_part_kernel<<<grid, 256, 0, cx.stream>>>(n, p, q, chunk_rows, A, lda, Aprev, ldap, B, ldb, ws);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
    reduce_parts(cx, p, q, (int)nch, ws, p, (int64_t)p * q, C, ldc, alpha, beta);
}

// ------------------------------------------------------------------------------------------------
// tiny dense helpers (single CTA): C = alpha*op(A)*op(B) + beta*C for matrices up to 256 x 256
// ------------------------------------------------------------------------------------------------
// 32 x 32 output tile per CTA, operands staged through shared memory in k-chunks of 32 (one global round trip per chunk
// instead of a dependent chain of k scalar loads per thread): 32^3 in ~4 us instead of ~14 us — the projected right-hand
// sides of the DE flows launch hundreds of these per step.
__global__ void __launch_bounds__(256) small_gemm_kernel(int p, int q, int k, const double* __restrict__ A, int lda, int tA,
                                                         const double* __restrict__ B, int ldb, int tB, double* __restrict__ C, int ldc,
                                                         double alpha, double beta, int64_t sA = 0, int64_t sB = 0, int64_t sC = 0) {
    // strided batch along blockIdx.z (all strides 0 for a single product)
    A += (int64_t)blockIdx.z * sA; B += (int64_t)blockIdx.z * sB; C += (int64_t)blockIdx.z * sC;
    __shared__ double As[32][33];   // As[l][i] = op(A)[i0 + i][l0 + l]
    __shared__ double Bs[32][33];   // Bs[l][j] = op(B)[l0 + l][j0 + j]
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int ti = threadIdx.x & 31, tj = threadIdx.x >> 5;   // thread -> rows ti, columns tj + 8 s
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int l0 = 0; l0 < k; l0 += 32) {
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int x = threadIdx.x & 31, y = (threadIdx.x >> 5) + 8 * s;
            {   // op(A)[i][l]: element (x, y) of the tile in the operand's fast direction
                const int i = tA ? i0 + y : i0 + x, l = tA ? l0 + x : l0 + y;
                double v = 0.0;
                if (i < p && l < k) v = tA ? A[l + (int64_t)i * lda] : A[i + (int64_t)l * lda];
                As[l - l0][i - i0] = v;
            }
            {
                const int l = tB ? l0 + y : l0 + x, j = tB ? j0 + x : j0 + y;
                double v = 0.0;
                if (l < k && j < q) v = tB ? B[j + (int64_t)l * ldb] : B[l + (int64_t)j * ldb];
                Bs[l - l0][j - j0] = v;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
            const double a = As[l][ti];
#pragma unroll
            for (int s = 0; s < 4; ++s) acc[s] = fma(a, Bs[l][tj + 8 * s], acc[s]);
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int i = i0 + ti, j = j0 + tj + 8 * s;
        if (i < p && j < q) {
            double* dst = C + i + (int64_t)j * ldc;
            *dst = (beta == 0.0 ? 0.0 : beta * (*dst)) + alpha * acc[s];
        }
    }
}
// `batch` independent products with operand strides sA, sB, sC (doubles) in ONE launch
inline void small_gemm(Ctx& cx, int p, int q, int k, const double* A, int lda, bool tA, const double* B, int ldb, bool tB,
                       double* C, int ldc, double alpha, double beta, int batch = 1, int64_t sA = 0, int64_t sB = 0, int64_t sC = 0) {
    if (p <= 0 || q <= 0 || batch <= 0) return;
    dim3 grid((unsigned)cdiv(p, 32), (unsigned)cdiv(q, 32), (unsigned)batch);
    small_gemm_kernel<<<grid, 256, 0, cx.stream>>>(p, q, k, A, lda, tA ? 1 : 0, B, ldb, tB ? 1 : 0, C, ldc, alpha, beta, sA, sB, sC);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

// Core update of the BUG-type integrators in ONE single-CTA launch:  Out = M*S0*N' (+ Add)
// (unconventional.jl:154 `set_u!(SIntegrator, M*u.S*N')` followed by the S-step increment).  M: p x r, S0: r x r, N: q x r,
// Add/Out: p x q, all with leading dimension ld; T (p x r, dense) is scratch.  Out may alias S0 (S0 is only read in phase 1).
__global__ void __launch_bounds__(1024) core_update_kernel(int p, int q, int r, const double* __restrict__ M, const double* __restrict__ S0,
                                                           const double* __restrict__ N, const double* __restrict__ Add,
                                                           double* __restrict__ Out, int ld, double* __restrict__ T) {
    for (int e = threadIdx.x; e < p * r; e += blockDim.x) {
        const int i = e % p, k = e / p;
        double s = 0.0;
        for (int l = 0; l < r; ++l) s = fma(M[i + (int64_t)l * ld], S0[l + (int64_t)k * ld], s);
        T[e] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < p * q; e += blockDim.x) {
        const int i = e % p, j = e / p;
        double s = Add ? Add[i + (int64_t)j * ld] : 0.0;
        double t = 0.0;
        for (int k = 0; k < r; ++k) t = fma(T[i + (int64_t)k * p], N[j + (int64_t)k * ld], t);
        Out[i + (int64_t)j * ld] = t + s;
    }
}
inline void core_update(Ctx& cx, int p, int q, int r, const double* M, const double* S0, const double* N, const double* Add, double* Out,
                        int ld, double* T) {
    core_update_kernel<<<1, 1024, 0, cx.stream>>>(p, q, r, M, S0, N, Add, Out, ld, T);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

// dst[rows x cols] (ldd) = alpha * op(src) (+ beta*dst)   — copies, transposes, scaled adds of small/tall matrices
__global__ void copy_mat_kernel(int64_t rows, int cols, const double* __restrict__ src, int64_t lds, int trans,
                                double* __restrict__ dst, int64_t ldd, double alpha, double beta) {
    int64_t tot = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e % rows;
        int64_t j = e / rows;
        double v = trans ? src[j + i * lds] : src[i + j * lds];
        double* d = dst + i + j * ldd;
        *d = alpha * v + (beta == 0.0 ? 0.0 : beta * (*d));
    }
}
inline void copy_mat(Ctx& cx, int64_t rows, int cols, const double* src, int64_t lds, bool trans, double* dst, int64_t ldd,
                     double alpha = 1.0, double beta = 0.0) {
    if (rows <= 0 || cols <= 0) return;
    int64_t tot = rows * cols;
    int blocks = (int)std::min<int64_t>(cdiv(tot, 256), 65535);
    copy_mat_kernel<<<blocks, 256, 0, cx.stream>>>(rows, cols, src, lds, trans ? 1 : 0, dst, ldd, alpha, beta);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

__global__ void fill_kernel(int64_t rows, int cols, double* __restrict__ dst, int64_t ldd, double v, double diag) {
    int64_t tot = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e % rows, j = e / rows;
        dst[i + j * ldd] = (i == j) ? diag : v;
    }
}
inline void fill_mat(Ctx& cx, int64_t rows, int cols, double* dst, int64_t ldd, double v, double diag) {
    if (rows <= 0 || cols <= 0) return;
    int blocks = (int)std::min<int64_t>(cdiv(rows * cols, 256), 65535);
    fill_kernel<<<blocks, 256, 0, cx.stream>>>(rows, cols, dst, ldd, v, diag);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

// Q[:, j] *= sign(R[j, j])  (R[j, j] == 0 keeps the column): turns the Q of a Householder QR of an (almost) orthonormal matrix
// back into that matrix, see ortho_complete in engine.cu
__global__ void sign_fix_cols_kernel(int rows, int cols, double* __restrict__ Q, int64_t ldq, const double* __restrict__ R, int64_t ldr) {
    const int tot = rows * cols;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += gridDim.x * blockDim.x) {
        const int i = e % rows, j = e / rows;
        if (R[j + (int64_t)j * ldr] < 0.0) Q[i + (int64_t)j * ldq] = -Q[i + (int64_t)j * ldq];
    }
}
inline void sign_fix_cols(Ctx& cx, int rows, int cols, double* Q, int64_t ldq, const double* R, int64_t ldr) {
    if (rows <= 0 || cols <= 0) return;
    sign_fix_cols_kernel<<<(unsigned)std::min<int64_t>(cdiv((int64_t)rows * cols, 256), 1024), 256, 0, cx.stream>>>(rows, cols, Q, ldq, R, ldr);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// sum_{i,j} (X[i,:]·W[j,:] - Yref[i,j])^2 and sum Yref^2  (X = U*S n x r, W = V m x r) -> out[2] partial per block
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) recon_err_kernel(int64_t n, int64_t m, int r, const double* __restrict__ X, int64_t ldx,
                                                        const double* __restrict__ W, int64_t ldw,
                                                        const double* __restrict__ Y, int64_t ldy, double* __restrict__ part) {
    constexpr int JT = 8;
    extern __shared__ double Ws[];  // [r][JT]
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int64_t j_begin = (int64_t)blockIdx.y * 512, j_end = min(m, j_begin + 512);
    double e2 = 0.0, y2 = 0.0;
    for (int64_t j0 = j_begin; j0 < j_end; j0 += JT) {
        __syncthreads();
        for (int e = threadIdx.x; e < r * JT; e += 128) {
            int c = e / JT, jj = e % JT;
            Ws[e] = (j0 + jj < j_end) ? W[(j0 + jj) + (int64_t)c * ldw] : 0.0;
        }
        __syncthreads();
        if (i < n) {
            double acc[JT];
#pragma unroll
            for (int jj = 0; jj < JT; ++jj) acc[jj] = 0.0;
            for (int c = 0; c < r; ++c) {
                double x = X[i + (int64_t)c * ldx];
#pragma unroll
                for (int jj = 0; jj < JT; ++jj) acc[jj] = fma(x, Ws[c * JT + jj], acc[jj]);
            }
#pragma unroll
            for (int jj = 0; jj < JT; ++jj)
                if (j0 + jj < j_end) {
                    double y = Y[i + (j0 + jj) * ldy];
                    double d = acc[jj] - y;
                    e2 = fma(d, d, e2);
                    y2 = fma(y, y, y2);
                }
        }
    }
    e2 = warp_sum(e2);
    y2 = warp_sum(y2);
    __shared__ double red[2][4];
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][w] = e2; red[1][w] = y2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t b = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        part[2 * b] = red[0][0] + red[0][1] + red[0][2] + red[0][3];
        part[2 * b + 1] = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    }
}
__global__ void sum_pairs_kernel(int64_t nblocks, const double* __restrict__ part, double* __restrict__ out) {
    double a = 0, b = 0;
    for (int64_t i = threadIdx.x; i < nblocks; i += blockDim.x) { a += part[2 * i]; b += part[2 * i + 1]; }
    a = warp_sum(a); b = warp_sum(b);
    __shared__ double ra[32], rb[32];
    if ((threadIdx.x & 31) == 0) { ra[threadIdx.x >> 5] = a; rb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0, sb = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sa += ra[w]; sb += rb[w]; }
        out[0] = sa; out[1] = sb;
    }
}

}  // namespace dlra
