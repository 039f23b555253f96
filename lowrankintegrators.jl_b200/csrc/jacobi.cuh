// One-block one-sided (Hestenes) Jacobi SVD of the small core matrix + tolerance-based rank selection.
// Replaces `U, S, V = svd(SIntegrator.u)` (LAPACK dgesdd) and
// `r_new = min(r_max, LowRankArithmetic.truncate_to_tolerance(S, tol))`
// at rank_adaptive_unconventional.jl:222-223.
//
// A (N x N) is copied to G, V = I; round-robin sweeps rotate column pairs of G (and V) until all pairs are
// orthogonal to working precision: G = A*V = P*diag(sigma)  =>  A = P * diag(sigma) * V'.
// One CTA of 32 warps, one warp per column pair and round; N <= 112 runs entirely in shared memory, larger
// cores (2r up to 256) run out of an L2-resident global scratch.  Singular values are sorted descending
// like LAPACK's; the rank rule restates oracle/dlra_oracle.py::truncate_to_tolerance exactly.
#pragma once
#include "common.cuh"

namespace dlra {

constexpr int JACOBI_SMEM_MAX_N = 112;

__global__ void __launch_bounds__(1024) jacobi_svd_kernel(int N, const double* __restrict__ A, int lda, double* __restrict__ Gws,
                                                          double* __restrict__ Vws, int use_smem, double* __restrict__ P, int ldp,
                                                          double* __restrict__ sigma, double* __restrict__ Q, int ldq,
                                                          double tol, int rcap, int* __restrict__ r_new_dev, int* __restrict__ r_new_host,
                                                          int max_sweeps) {
    extern __shared__ __align__(16) double jsm[];
    __shared__ double s_off[32];
    __shared__ double s_sig[256];
    __shared__ int s_rank[256];
    __shared__ int s_continue;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int ld = N;
    double* G = use_smem ? jsm : Gws;
    double* V = use_smem ? jsm + (size_t)N * N : Vws;
    for (int e = tid; e < N * N; e += blockDim.x) {
        int i = e % N, j = e / N;
        G[i + j * ld] = A[i + (int64_t)j * lda];
        V[i + j * ld] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int Np = (N + 1) & ~1;
    const double eps_rot = 1e-15;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        double off_max = 0.0;
        for (int round = 0; round < Np - 1; ++round) {
            for (int pi = warp; pi < Np / 2; pi += nwarps) {
                int p, q;
                if (pi == 0) { p = Np - 1; q = round; }
                else { p = (round + pi) % (Np - 1); q = (round - pi + (Np - 1)) % (Np - 1); }
                if (p >= N || q >= N) continue;
                if (p > q) { int t = p; p = q; q = t; }
                double* gp = G + p * ld; double* gq = G + q * ld;
                double a = 0, b = 0, g = 0;
                for (int i = lane; i < N; i += 32) {
                    double x = gp[i], y = gq[i];
                    a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
                }
                a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
                const double denom = sqrt(a * b);
                if (denom > 0.0 && fabs(g) > eps_rot * denom) {
                    off_max = fmax(off_max, fabs(g) / denom);
                    const double zeta = (b - a) / (2.0 * g);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    double* vp = V + p * ld; double* vq = V + q * ld;
                    for (int i = lane; i < N; i += 32) {
                        double x = gp[i], y = gq[i];
                        gp[i] = c * x - s * y; gq[i] = s * x + c * y;
                        double u = vp[i], w = vq[i];
                        vp[i] = c * u - s * w; vq[i] = s * u + c * w;
                    }
                }
            }
            __syncthreads();
        }
        off_max = warp_max(off_max);
        if (lane == 0) s_off[warp] = off_max;
        __syncthreads();
        if (tid == 0) {
            double mx = 0.0;
            for (int w = 0; w < nwarps; ++w) mx = fmax(mx, s_off[w]);
            s_continue = (mx > 0.0) ? 1 : 0;   // a sweep without any rotation => converged
        }
        __syncthreads();
        if (!s_continue) break;
    }
    // column norms -> singular values
    for (int j = warp; j < N; j += nwarps) {
        double a = 0;
        for (int i = lane; i < N; i += 32) { double x = G[i + j * ld]; a = fma(x, x, a); }
        a = warp_sum(a);
        if (lane == 0) s_sig[j] = sqrt(a);
    }
    __syncthreads();
    for (int j = tid; j < N; j += blockDim.x) {
        const double sj = s_sig[j];
        int rk = 0;
        for (int k = 0; k < N; ++k) { double sk = s_sig[k]; rk += (sk > sj || (sk == sj && k < j)) ? 1 : 0; }
        s_rank[j] = rk;
        sigma[rk] = sj;
    }
    __syncthreads();
    for (int e = tid; e < N * N; e += blockDim.x) {
        int i = e % N, j = e / N;
        const double sj = s_sig[j];
        const int rk = s_rank[j];
        P[i + (int64_t)rk * ldp] = (sj > 0.0) ? G[i + j * ld] / sj : 0.0;
        Q[i + (int64_t)rk * ldq] = V[i + j * ld];
    }
    __syncthreads();
    if (tid == 0 && r_new_dev) {
        // truncate_to_tolerance (UNVERIFIED third-party semantics, see oracle): accumulate from the tail
        double s = 0.0;
        int r = N;
        for (int k = N - 1; k >= 0; --k) {
            const double sg = sigma[k];
            s += sg * sg;
            if (s > tol * tol) break;
            r -= 1;
        }
        if (r > rcap) r = rcap;
        if (r < 1) r = 1;
        *r_new_dev = r;
        if (r_new_host) *r_new_host = r;
    }
}

inline size_t jacobi_ws_doubles(int N) { return (size_t)2 * N * N; }

inline void jacobi_svd(Ctx& cx, int N, const double* A, int lda, double* ws, double* P, int ldp, double* sigma, double* Q, int ldq,
                       double tol, int rcap, int* r_new_dev, int* r_new_host) {
    DLRA_REQUIRE(N >= 1 && N <= 256, "core SVD supports 1 <= N <= 256");
    const int use_smem = (N <= JACOBI_SMEM_MAX_N) ? 1 : 0;
    const size_t smem = use_smem ? (size_t)2 * N * N * sizeof(double) : 0;
    static bool attr_set = false;
    if (!attr_set) {
        DLRA_CUDA(cudaFuncSetAttribute(jacobi_svd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)2 * JACOBI_SMEM_MAX_N * JACOBI_SMEM_MAX_N * sizeof(double))));
        attr_set = true;
    }
    int threads = 1024;
    if (N <= 32) threads = 512;
    jacobi_svd_kernel<<<1, threads, smem, cx.stream>>>(N, A, lda, ws, ws + (size_t)N * N, use_smem, P, ldp, sigma, Q, ldq, tol, rcap,
                                                        r_new_dev, r_new_host, 60);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

}  // namespace dlra
