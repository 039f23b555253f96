// MatrixDEProblem sub-flows on the device.
// Replaces the default projected right-hand sides of the reference
//     K_rhs(K, V, t) = Matrix( f(TwoFactorRepresentation(K, V), t) * V )
//     L_rhs(L, U, t) = Matrix( f(TwoFactorRepresentation(U, L), t)' * U )
//     S_rhs(S,(U,V),t) = -/+ Matrix( U' * f(SVDLikeRepresentation(U,S,V), t) * V )
// (projector_splitting.jl:52-80, unconventional.jl:51-79, rank_adaptive_unconventional.jl:59-86) and the
// OrdinaryDiffEq sub-integrators `set_u!(I,u); step!(I, dt, true); I.u` for the device-evaluable family
//     F(X,t) = A·X + X·Bᵀ + G·Hᵀ + c·(D1·X) .* (D2·X).
// Like the reference's LowRankArithmetic evaluation nothing n x m is ever formed: with V (resp. U) orthonormal
//   K-flow:  F(K Vᵀ)V   = A·K + K·(VᵀBᵀV) + G·(HᵀV) + c·Σ_ab (D1K)[:,a].*(D2K)[:,b]·T_V[a,b,:]
//   L-flow:  F(U Lᵀ)ᵀU  = B·L + L·(UᵀAᵀU) + H·(GᵀU) + c·Σ_ab L[:,a].*L[:,b]·T_U[a,b,:]
//   S-flow:  UᵀF(USVᵀ)V = (UᵀAU)·S + S·(VᵀBᵀV) + (UᵀG)(HᵀV) + c·Σ TU[a',b',:]·S[a',a]·S[b',b]·T_V[a,b,:]
// with the third-order tensors T_V[a,b,c] = Σ_j V[j,a]V[j,b]V[j,c] and T_U[a,b,c] = Σ_i (D1U)[i,a](D2U)[i,b]U[i,c]
// formed once per flow (the Khatri-Rao structure LowRankArithmetic's `.*` creates).
// The explicit RK schemes (Euler, RK4, Tsit5 fixed/adaptive) restate oracle/dlra_oracle.py::ode_advance 1:1.
#pragma once
#include "engine.cuh"
#include <mutex>

namespace dlra {

// ------------------------------------------------------------------------------------------------------
// elementwise / reduction helpers
// ------------------------------------------------------------------------------------------------------
struct LinComb {
    const double* k[8];
    double c[8];
    int nk;
};
// out = base + Σ c_j k_j   (base may be null)
__global__ void lincomb_kernel(int64_t n, const double* __restrict__ base, LinComb lc, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = base ? base[i] : 0.0;
#pragma unroll 8
        for (int j = 0; j < lc.nk; ++j) s = fma(lc.c[j], lc.k[j][i], s);
        out[i] = s;
    }
}
inline void lincomb(Ctx& cx, int64_t n, const double* base, const LinComb& lc, double* out) {
    int blocks = (int)std::min<int64_t>(cdiv(n, 256), 4 * 148);
    lincomb_kernel<<<blocks, 256, 0, cx.stream>>>(n, base, lc, out);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

// part[b] = Σ (x_i / (abstol + max(|u_i|,|v_i|)·reltol))²
__global__ void __launch_bounds__(256) scaled_sq_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ u,
                                                        const double* __restrict__ v, double abstol, double reltol,
                                                        double* __restrict__ part) {
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double sk = abstol + fmax(fabs(u[i]), fabs(v[i])) * reltol;
        const double q = x[i] / sk;
        s = fma(q, q, s);
    }
    s = warp_sum(s);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        part[blockIdx.x] = t;
    }
}
__global__ void sum_parts_kernel(int nb, const double* __restrict__ part, double* __restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += 32) s += part[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) *out = s;
}

// ------------------------------------------------------------------------------------------------------
// operator application  Y (rows x cols, ldy) = beta*Y + alpha * Op * X      (Op: rows x rows)
// ------------------------------------------------------------------------------------------------------
__global__ void spmm_csr_kernel(int64_t rows, int cols, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                const double* __restrict__ vals, const double* __restrict__ X, int64_t ldx,
                                double* __restrict__ Y, int64_t ldy, double alpha, double beta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= rows) return;
    double s = 0.0;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) s = fma(vals[k], X[colind[k] + (int64_t)c * ldx], s);
    double* y = Y + i + (int64_t)c * ldy;
    *y = (beta == 0.0 ? 0.0 : beta * (*y)) + alpha * s;
}

inline void apply_op(Ctx& cx, const dlra_operator& op, int64_t rows, int cols, const double* X, int64_t ldx, double* Y, int64_t ldy,
                     double alpha, double beta) {
    switch (op.kind) {
        case DLRA_OP_DENSE:
            gemm_nn(cx, rows, (int)op.cols, cols, op.dense, op.ld, nullptr, 0, X, ldx, false, Y, ldy, alpha * op.scale, beta);
            break;
        case DLRA_OP_CSR: {
            dim3 grid((unsigned)cdiv(rows, 128), (unsigned)cols);
            spmm_csr_kernel<<<grid, 128, 0, cx.stream>>>(rows, cols, op.rowptr, op.colind, op.values, X, ldx, Y, ldy, alpha * op.scale, beta);
            cx.launches++;
            DLRA_CUDA(cudaGetLastError());
            break;
        }
        case DLRA_OP_IDENTITY_SCALED:
            copy_mat(cx, rows, cols, X, ldx, false, Y, ldy, alpha * op.scale, beta);
            break;
        default:
            if (beta == 0.0) fill_mat(cx, rows, cols, Y, ldy, 0.0, 0.0);
            break;
    }
}

// ------------------------------------------------------------------------------------------------------
// third-order tensors and the Hadamard term
// ------------------------------------------------------------------------------------------------------
// Tpart[chunk][a + p*(b + p*c)] = Σ_{i in chunk} P[i,a] Q[i,b] W[i,c]   (P,Q,W: rows x p)
__global__ void __launch_bounds__(256) tensor3_part_kernel(int64_t rows, int p, int64_t chunk_rows, const double* __restrict__ P, int64_t ldp,
                                                          const double* __restrict__ Q, int64_t ldq, const double* __restrict__ Wm,
                                                          int64_t ldw, double* __restrict__ Tpart) {
    extern __shared__ double sm3[];   // three tiles [p][32]
    double* Ps = sm3; double* Qs = Ps + (size_t)p * 32; double* Ws = Qs + (size_t)p * 32;
    const int64_t r0 = (int64_t)blockIdx.x * chunk_rows, r1 = min(rows, r0 + chunk_rows);
    const int64_t p3 = (int64_t)p * p * p;
    // each thread owns entries e = tid + 256*l of the slab assigned to blockIdx.y (slab = 4096 entries)
    const int64_t e0 = (int64_t)blockIdx.y * 4096;
    double acc[16];
#pragma unroll
    for (int l = 0; l < 16; ++l) acc[l] = 0.0;
    for (int64_t rr = r0; rr < r1; rr += 32) {
        __syncthreads();
        for (int e = threadIdx.x; e < p * 32; e += 256) {
            const int a = e / 32, i = e % 32;
            const bool ok = rr + i < r1;
            Ps[e] = ok ? P[rr + i + (int64_t)a * ldp] : 0.0;
            Qs[e] = ok ? Q[rr + i + (int64_t)a * ldq] : 0.0;
            Ws[e] = ok ? Wm[rr + i + (int64_t)a * ldw] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            const int64_t e = e0 + threadIdx.x + 256 * l;
            if (e < p3) {
                const int a = (int)(e % p), b = (int)((e / p) % p), c = (int)(e / ((int64_t)p * p));
                double s = acc[l];
                for (int i = 0; i < 32; ++i) s = fma(Ps[a * 32 + i] * Qs[b * 32 + i], Ws[c * 32 + i], s);
                acc[l] = s;
            }
        }
    }
#pragma unroll
    for (int l = 0; l < 16; ++l) {
        const int64_t e = e0 + threadIdx.x + 256 * l;
        if (e < p3) Tpart[(int64_t)blockIdx.x * p3 + e] = acc[l];
    }
}


// The same third-order tensor on the fp64 tensor pipe:  T[(a,b), c] = Σ_i Z[i,(a,b)]·W[i,c]  with  Z[i,(a,b)] = P[i,a]·Q[i,b]  formed in
// registers — a (p² x rows)·(rows x p) product.  A CTA owns 128 (a,b) pairs (16 per warp: two 8-row DMMA blocks) x all c and a chunk of rows;
// 32-row tiles of P, Q, W are staged row-major with a stride of 68 doubles (conflict-free fragment loads).  20x the scalar kernel above at
// p = 32 (profiles/r02): that kernel spends three shared-memory loads per multiply-add.
constexpr int T3_LD = 68, T3_ROWS = 32;
constexpr int T3_SMEM_BYTES = 3 * T3_ROWS * T3_LD * (int)sizeof(double);
template <int NB>   // NB = ceil(p / 8) column blocks of c (4 for p <= 32, 8 for p <= 64)
__global__ void __launch_bounds__(256) tensor3_dmma_kernel(int64_t rows, int p, int64_t chunk_rows, const double* __restrict__ P, int64_t ldp,
                                                          const double* __restrict__ Q, int64_t ldq, const double* __restrict__ Wm,
                                                          int64_t ldw, double* __restrict__ Tpart) {
    extern __shared__ __align__(16) double t3s[];
    double* Ps = t3s; double* Qs = Ps + T3_ROWS * T3_LD; double* Ws = Qs + T3_ROWS * T3_LD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, k = lane & 3;
    const int64_t r0 = (int64_t)blockIdx.x * chunk_rows, r1 = min(rows, r0 + chunk_rows);
    const int p2 = p * p;
    const int mbase = blockIdx.y * 128 + warp * 16;
    int ia[2], ib[2]; bool mok[2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        const int m = mbase + 8 * mb + g;
        mok[mb] = m < p2;
        ia[mb] = mok[mb] ? m % p : 0;
        ib[mb] = mok[mb] ? m / p : 0;
    }
    double acc[2][NB][2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) { acc[mb][nb][0] = 0.0; acc[mb][nb][1] = 0.0; }
    for (int64_t rr = r0; rr < r1; rr += T3_ROWS) {
        __syncthreads();
        // stage [32 rows][p (+ zero padding up to 8·NB)] of the three operands: global reads run along the rows of a column
        for (int e = threadIdx.x; e < 8 * NB * T3_ROWS; e += 256) {
            const int c = e / T3_ROWS, i = e % T3_ROWS;
            const bool ok = (rr + i < r1) && (c < p);
            Ps[i * T3_LD + c] = ok ? P[rr + i + (int64_t)c * ldp] : 0.0;
            Qs[i * T3_LD + c] = ok ? Q[rr + i + (int64_t)c * ldq] : 0.0;
            Ws[i * T3_LD + c] = ok ? Wm[rr + i + (int64_t)c * ldw] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < T3_ROWS / 4; ++ks) {
            const int row = (4 * ks + k) * T3_LD;
            double af[2], bf[NB];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) af[mb] = mok[mb] ? Ps[row + ia[mb]] * Qs[row + ib[mb]] : 0.0;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) bf[nb] = Ws[row + 8 * nb + g];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], af[mb], bf[nb]);
        }
    }
    double* out = Tpart + (int64_t)blockIdx.x * p2 * p;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        const int m = mbase + 8 * mb + g;
        if (m >= p2) continue;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * nb + 2 * k + e;
                if (c < p) out[m + (int64_t)p2 * c] = acc[mb][nb][e];
            }
    }
}

// out[i, c] (+)= coef * Σ_{a,b} P[i,a] Q[i,b] T[a + p*(b + p*c)]      (rows x p operands, p outputs per row)
__global__ void __launch_bounds__(128) hadamard_rows_kernel(int64_t rows, int p, const double* __restrict__ P, int64_t ldp,
                                                           const double* __restrict__ Q, int64_t ldq, const double* __restrict__ T,
                                                           double* __restrict__ out, int64_t ldo, double coef) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= rows) return;
    const double* Tc = T + (int64_t)c * p * p;
    double s = 0.0;
    for (int b = 0; b < p; ++b) {
        const double qb = Q[i + (int64_t)b * ldq];
        double sb = 0.0;
        for (int a = 0; a < p; ++a) sb = fma(P[i + (int64_t)a * ldp], Tc[a + b * p], sb);
        s = fma(qb, sb, s);
    }
    out[i + (int64_t)c * ldo] += coef * s;
}


// Same contraction for p <= 32 with the row of P and all p outputs of a row held in registers: the p x p slab T[:, b, :] is
// staged in shared memory once per b and read as warp-wide broadcasts (one LDS.128 per two FMAs), so the kernel runs close
// to the DFMA rate instead of issuing one cached global load per FMA (10x on the Burgers flows, profiles/r02).
__global__ void __launch_bounds__(64) hadamard_rows32_kernel(int64_t rows, int p, const double* __restrict__ P, int64_t ldp,
                                                            const double* __restrict__ Q, int64_t ldq, const double* __restrict__ T,
                                                            double* __restrict__ out, int64_t ldo, double coef) {
    __shared__ __align__(16) double Ts[32][32];   // Ts[a][c] = T[a, b, c]
    const int64_t i = (int64_t)blockIdx.x * 64 + threadIdx.x;
    const bool ok = i < rows;
    double pr[32], acc[32];
#pragma unroll
    for (int a = 0; a < 32; ++a) { pr[a] = (ok && a < p) ? P[i + (int64_t)a * ldp] : 0.0; acc[a] = 0.0; }
    for (int b = 0; b < p; ++b) {
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * 32; e += 64) {
            const int a = e & 31, c = e >> 5;
            Ts[a][c] = (a < p && c < p) ? T[a + (int64_t)p * (b + (int64_t)p * c)] : 0.0;
        }
        __syncthreads();
        const double qb = ok ? Q[i + (int64_t)b * ldq] : 0.0;
        double t2[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) t2[c] = 0.0;
#pragma unroll
        for (int a = 0; a < 32; ++a) {
            const double pa = pr[a];
            const double2* row = reinterpret_cast<const double2*>(&Ts[a][0]);
#pragma unroll
            for (int c2 = 0; c2 < 16; ++c2) {
                const double2 tv = row[c2];
                t2[2 * c2] = fma(pa, tv.x, t2[2 * c2]);
                t2[2 * c2 + 1] = fma(pa, tv.y, t2[2 * c2 + 1]);
            }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = fma(qb, t2[c], acc[c]);
    }
    if (ok) {
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (c < p) out[i + (int64_t)c * ldo] += coef * acc[c];
    }
}

// ------------------------------------------------------------------------------------------------------
// per-engine DE workspace
// ------------------------------------------------------------------------------------------------------
struct DeWork {
    DevBuf stages;      // 10 state-sized buffers for the RK schemes
    DevBuf nbuf[4];     // n x W scratch: A·U / D1·X / D2·X / X·(VᵀB_kᵀV)
    DevBuf mbuf[2];     // m x W scratch: B·V / X·(UᵀA_kᵀU)
    DevBuf tsm;         // projected r x r matrices of the two-sided terms
    DevBuf tens[3];     // T_V, T_U, contraction scratch
    DevBuf tpart;       // partial tensors
    DevBuf small;       // r x r matrices
    DevBuf red;         // reduction partials + scalar
    double* scal_host = nullptr;
};

inline DeWork* de_work(dlra_engine* e);

}  // namespace dlra

namespace dlra {

static std::vector<std::pair<dlra_engine*, DeWork*>>& de_registry() {
    static std::vector<std::pair<dlra_engine*, DeWork*>> r;
    return r;
}
// handles may be driven from different host threads (one handle per thread): the registry itself is shared
static std::mutex& de_registry_mutex() {
    static std::mutex mu;
    return mu;
}
inline DeWork* de_work(dlra_engine* e) {
    std::lock_guard<std::mutex> lock(de_registry_mutex());
    for (auto& pr : de_registry()) if (pr.first == e) return pr.second;
    DeWork* w = new DeWork();
    DLRA_CUDA(cudaHostAlloc(&w->scal_host, 4 * sizeof(double), cudaHostAllocDefault));
    de_registry().emplace_back(e, w);
    return w;
}
inline void de_release(dlra_engine* e) {
    std::lock_guard<std::mutex> lock(de_registry_mutex());
    auto& reg = de_registry();
    for (size_t i = 0; i < reg.size(); ++i)
        if (reg[i].first == e) {
            DeWork* w = reg[i].second;
            w->stages.release();
            for (auto& b : w->nbuf) b.release();
            for (auto& b : w->mbuf) b.release();
            for (auto& b : w->tens) b.release();
            w->tpart.release(); w->small.release(); w->red.release(); w->tsm.release();
            if (w->scal_host) cudaFreeHost(w->scal_host);
            delete w;
            reg.erase(reg.begin() + i);
            return;
        }
}
inline void de_rank_changed(dlra_engine* e) {
    // alg_recache re-inits the ODE integrators (rank_adaptive_unconventional.jl:150-164): controller state starts afresh
    for (int f = 0; f < 3; ++f) { e->sub[f].dt_next = -1.0; e->sub[f].qold = 1e-4; }
}

inline void de_rhs_set(dlra_engine* e, const dlra_operator* A, const dlra_operator* B, const double* G, int64_t ldg, const double* H,
                       int64_t ldh, int q, const dlra_operator* D1, const dlra_operator* D2, double c_had) {
    DLRA_REQUIRE(e->comm.nranks == 1, "DE right-hand sides are single-GPU (operator apply would need a halo/all-gather of K)");
    auto chk = [&](const dlra_operator* op, int64_t dim, const char* nm) {
        if (!op || op->kind == DLRA_OP_NONE) return;
        DLRA_REQUIRE(op->kind == DLRA_OP_DENSE || op->kind == DLRA_OP_CSR || op->kind == DLRA_OP_IDENTITY_SCALED, "unknown operator kind");
        if (op->kind != DLRA_OP_IDENTITY_SCALED) DLRA_REQUIRE(op->rows == dim && op->cols == dim, std::string("operator has the wrong shape: ") + nm);
        if (op->kind == DLRA_OP_DENSE) DLRA_REQUIRE(op->dense && op->ld >= dim, "dense operator pointer / ld");
        if (op->kind == DLRA_OP_CSR) DLRA_REQUIRE(op->rowptr && op->colind && op->values, "CSR operator pointers");
    };
    chk(A, e->n, "A"); chk(B, e->m, "B"); chk(D1, e->n, "D1"); chk(D2, e->n, "D2");
    DLRA_REQUIRE(q >= 0 && (q == 0 || (G && H && ldg >= e->n && ldh >= e->m)), "forcing factors G (n x q), H (m x q)");
    RhsCfg r;
    r.set = true;
    if (A) r.A = *A; if (B) r.B = *B; if (D1) r.D1 = *D1; if (D2) r.D2 = *D2;
    r.G = G; r.ldg = ldg; r.H = H; r.ldh = ldh; r.q = q; r.c_had = c_had;
    if (c_had != 0.0) DLRA_REQUIRE(r.D1.kind != DLRA_OP_NONE && r.D2.kind != DLRA_OP_NONE, "Hadamard term needs D1 and D2");
    e->rhs = r;
    de_work(e);
}

// one more term A_k·X·B_kᵀ of the right-hand side (operators as in de_rhs_set; identity-scaled allowed on either side)
inline void de_rhs_add_term(dlra_engine* e, const dlra_operator* A, const dlra_operator* B) {
    DLRA_REQUIRE(e->rhs.set, "install the right-hand side first (dlra_rhs_set; all of its terms may be empty)");
    DLRA_REQUIRE(A && B, "both operators of a two-sided term are needed (use DLRA_OP_IDENTITY_SCALED for the identity)");
    auto chk = [&](const dlra_operator* op, int64_t dim) {
        DLRA_REQUIRE(op->kind == DLRA_OP_DENSE || op->kind == DLRA_OP_CSR || op->kind == DLRA_OP_IDENTITY_SCALED, "unknown operator kind");
        if (op->kind != DLRA_OP_IDENTITY_SCALED) DLRA_REQUIRE(op->rows == dim && op->cols == dim, "operator of a two-sided term has the wrong shape");
        if (op->kind == DLRA_OP_DENSE) DLRA_REQUIRE(op->dense && op->ld >= dim, "dense operator pointer / ld");
        if (op->kind == DLRA_OP_CSR) DLRA_REQUIRE(op->rowptr && op->colind && op->values, "CSR operator pointers");
    };
    chk(A, e->n); chk(B, e->m);
    DLRA_REQUIRE(e->rhs.terms.size() < 64, "at most 64 two-sided terms");
    e->rhs.terms.emplace_back(*A, *B);
}

// T (p^3) = Σ_rows P .* Q .* W triple products, deterministic two-stage reduction
inline void tensor3(dlra_engine* e, DeWork* w, int64_t rows, int p, const double* P, int64_t ldp, const double* Q, int64_t ldq,
                    const double* Wm, int64_t ldw, double* T) {
    Ctx& cx = e->cx;
    DLRA_REQUIRE(p <= 64, "the Hadamard (column-wise nonlinear) term supports factor widths up to 64");
    const int64_t p3 = (int64_t)p * p * p;
    static const bool legacy = getenv("DLRA_TENSOR3_LEGACY") != nullptr;
    int64_t nch;
    if (!legacy) {
        const int slabs = (int)cdiv((int64_t)p * p, 128);
        const int64_t want_chunks = std::max<int64_t>(1, (2 * cx.num_sms) / slabs);
        const int64_t chunk_rows = round_up(cdiv(rows, want_chunks), T3_ROWS);
        nch = cdiv(rows, chunk_rows);
        w->tpart.ensure(nch * p3, cx.stream);
        static unsigned long long attr_devs = 0;
        if (first_use_on_this_device(attr_devs)) {
            DLRA_CUDA(cudaFuncSetAttribute(tensor3_dmma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM_BYTES));
            DLRA_CUDA(cudaFuncSetAttribute(tensor3_dmma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM_BYTES));
        }
        dim3 grid((unsigned)nch, (unsigned)slabs);
        if (p <= 32) tensor3_dmma_kernel<4><<<grid, 256, T3_SMEM_BYTES, cx.stream>>>(rows, p, chunk_rows, P, ldp, Q, ldq, Wm, ldw, w->tpart.p);
        else tensor3_dmma_kernel<8><<<grid, 256, T3_SMEM_BYTES, cx.stream>>>(rows, p, chunk_rows, P, ldp, Q, ldq, Wm, ldw, w->tpart.p);
    } else {
        const int slabs = (int)cdiv(p3, 4096);
        int64_t want_chunks = std::max<int64_t>(1, (2 * cx.num_sms) / slabs);
        int64_t chunk_rows = round_up(cdiv(rows, want_chunks), 32);
        nch = cdiv(rows, chunk_rows);
        w->tpart.ensure(nch * p3, cx.stream);
        dim3 grid((unsigned)nch, (unsigned)slabs);
        tensor3_part_kernel<<<grid, 256, (size_t)3 * p * 32 * sizeof(double), cx.stream>>>(rows, p, chunk_rows, P, ldp, Q, ldq, Wm, ldw, w->tpart.p);
    }
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
    // view the p^3 entries as a (p^2) x p matrix for the generic fixed-order reduction
    reduce_parts(cx, p * p, p, (int)nch, w->tpart.p, (int64_t)p * p, p3, T, (int64_t)p * p, 1.0, 0.0);
}

// ------------------------------------------------------------------------------------------------------
// RK driver (restates oracle.ode_advance)
// ------------------------------------------------------------------------------------------------------
struct FlowRhs {
    virtual void eval(const double* X, double* out, double t) = 0;   // out = f(X, t), dense state of N doubles
    virtual ~FlowRhs() {}
};

static const double TS_C[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
static const double TS_A[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
static const double TS_BT[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                                0.5823571654525552, -0.45808210592918697, 0.015151515151515152};

inline double rms_scaled(dlra_engine* e, DeWork* w, int64_t N, const double* x, const double* u, const double* v, double abstol, double reltol) {
    Ctx& cx = e->cx;
    const int nb = (int)std::min<int64_t>(cdiv(N, 256), 2 * cx.num_sms);
    w->red.ensure(nb + 8, cx.stream);
    scaled_sq_kernel<<<nb, 256, 0, cx.stream>>>(N, x, u, v, abstol, reltol, w->red.p);
    sum_parts_kernel<<<1, 32, 0, cx.stream>>>(nb, w->red.p, w->red.p + nb);
    cx.launches += 2;
    DLRA_CUDA(cudaMemcpyAsync(w->scal_host, w->red.p + nb, sizeof(double), cudaMemcpyDeviceToHost, cx.stream));
    DLRA_CUDA(cudaStreamSynchronize(cx.stream));
    return sqrt(w->scal_host[0] / (double)N);
}

// one Tsit5 step from (u, k1): fills ks[1..6], unew
inline void tsit5_stages(dlra_engine* e, FlowRhs& f, int64_t N, const double* u, double t, double h, double** ks, double* utmp, double* unew) {
    for (int s = 1; s < 7; ++s) {
        LinComb lc;
        lc.nk = 0;
        for (int j = 0; j < s; ++j)
            if (TS_A[s][j] != 0.0) { lc.k[lc.nk] = ks[j]; lc.c[lc.nk] = h * TS_A[s][j]; lc.nk++; }
        double* dst = (s < 6) ? utmp : unew;
        lincomb(e->cx, N, u, lc, dst);
        f.eval(dst, ks[s], (s < 6) ? t + TS_C[s] * h : t + h);
    }
}

inline void ode_advance(dlra_engine* e, SubStepperCfg& st, FlowRhs& f, int64_t N, double* X, double t0, double dt,
                        FsalCarry* carry = nullptr) {
    NvtxRange nvtx_ode("dlra:ode_advance");
    DeWork* w = de_work(e);
    Ctx& cx = e->cx;
    w->stages.ensure(10 * N, cx.stream);
    double* ks[7];
    for (int i = 0; i < 7; ++i) ks[i] = w->stages.p + (int64_t)i * N;
    double* utmp = w->stages.p + 7 * N;
    double* unew = w->stages.p + 8 * N;
    double* etmp = w->stages.p + 9 * N;
    // first stage of the first sub-step: evaluated, or (carry) the derivative the previous call ended with
    auto first_stage = [&](double t) {
        if (carry && carry->valid && carry->N == N) {
            DLRA_CUDA(cudaMemcpyAsync(ks[0], carry->k.p, N * sizeof(double), cudaMemcpyDeviceToDevice, cx.stream));
        } else {
            f.eval(X, ks[0], t);
            st.nfev += 1;
        }
    };
    auto keep_stage = [&](const double* k) {
        carry->k.ensure(N, cx.stream);
        DLRA_CUDA(cudaMemcpyAsync(carry->k.p, k, N * sizeof(double), cudaMemcpyDeviceToDevice, cx.stream));
        carry->valid = true; carry->N = N;
    };
    if (st.ode != DLRA_ODE_TSIT5) {
        const double h = dt / st.nsub;
        double t = t0;
        for (int it = 0; it < st.nsub; ++it) {
            if (it == 0) first_stage(t);
            else { f.eval(X, ks[0], t); st.nfev += 1; }
            if (st.ode == DLRA_ODE_EULER) {
                LinComb lc; lc.nk = 1; lc.k[0] = ks[0]; lc.c[0] = h;
                lincomb(cx, N, X, lc, X);
            } else if (st.ode == DLRA_ODE_RK4) {
                LinComb a; a.nk = 1; a.k[0] = ks[0]; a.c[0] = 0.5 * h;
                lincomb(cx, N, X, a, utmp);
                f.eval(utmp, ks[1], t + 0.5 * h);
                a.k[0] = ks[1];
                lincomb(cx, N, X, a, utmp);
                f.eval(utmp, ks[2], t + 0.5 * h);
                a.k[0] = ks[2]; a.c[0] = h;
                lincomb(cx, N, X, a, utmp);
                f.eval(utmp, ks[3], t + h);
                LinComb b; b.nk = 4;
                b.k[0] = ks[0]; b.k[1] = ks[1]; b.k[2] = ks[2]; b.k[3] = ks[3];
                b.c[0] = h / 6.0; b.c[1] = h / 3.0; b.c[2] = h / 3.0; b.c[3] = h / 6.0;
                lincomb(cx, N, X, b, X);
                st.nfev += 3;
            } else {
                tsit5_stages(e, f, N, X, t, h, ks, utmp, unew);
                DLRA_CUDA(cudaMemcpyAsync(X, unew, N * sizeof(double), cudaMemcpyDeviceToDevice, cx.stream));
                st.nfev += 6;
            }
            t += h;
        }
        if (carry) {
            if (st.ode == DLRA_ODE_TSIT5_FIXED) keep_stage(ks[6]);
            else { f.eval(X, ks[0], t); st.nfev += 1; keep_stage(ks[0]); }
        }
        return;
    }
    // adaptive Tsit5 with the PI controller of SURVEY.md Appendix B
    const double tend = t0 + dt;
    double t = t0;
    first_stage(t);   // FSAL invalidated by set_u! (no carry) or kept (hybrid Z integrator)
    if (st.dt_next <= 0.0) {
        const double d0 = rms_scaled(e, w, N, X, X, X, st.abstol, st.reltol);
        const double d1 = rms_scaled(e, w, N, ks[0], X, X, st.abstol, st.reltol);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = std::min(h0, dt);
        LinComb a; a.nk = 1; a.k[0] = ks[0]; a.c[0] = h0;
        lincomb(cx, N, X, a, utmp);
        f.eval(utmp, ks[1], t + h0);
        st.nfev += 1;
        LinComb d; d.nk = 2; d.k[0] = ks[1]; d.c[0] = 1.0; d.k[1] = ks[0]; d.c[1] = -1.0;
        lincomb(cx, N, nullptr, d, etmp);
        const double d2 = rms_scaled(e, w, N, etmp, X, X, st.abstol, st.reltol) / h0;
        const double dm = std::max(d1, d2);
        const double h1 = (dm <= 1e-15) ? std::max(1e-6, h0 * 1e-3) : pow(0.01 / dm, 1.0 / 5.0);
        st.dt_next = std::min(std::min(100.0 * h0, h1), dt);
    }
    double h = st.dt_next;
    const double beta1 = 7.0 / 50.0, beta2 = 2.0 / 25.0, gamma = 0.9, qmin = 0.2, qmax = 10.0;
    int64_t guard = 0;
    while ((tend - t) > 1e-14 * std::max(1.0, fabs(tend))) {
        if (++guard > st.maxiters)
            throw CudaError(DLRA_EMAXITERS, "adaptive Tsit5 reached maxiters before the end of the step (stiff or unstable projected flow; "
                                            "factors and controller state are those before the step)");
        h = std::min(h, tend - t);
        tsit5_stages(e, f, N, X, t, h, ks, utmp, unew);
        st.nfev += 6;
        LinComb el; el.nk = 7;
        for (int j = 0; j < 7; ++j) { el.k[j] = ks[j]; el.c[j] = h * TS_BT[j]; }
        lincomb(cx, N, nullptr, el, etmp);
        const double EEst = rms_scaled(e, w, N, etmp, X, unew, st.abstol, st.reltol);
        DLRA_REQUIRE(std::isfinite(EEst), "non-finite state in the adaptive sub-stepper (diverged earlier step or bad input)");
        if (EEst <= 1.0) {
            const double q11 = pow(std::max(EEst, 1e-30), beta1);
            double q = q11 / pow(st.qold, beta2);
            q = std::max(1.0 / qmax, std::min(1.0 / qmin, q / gamma));
            st.qold = std::max(EEst, 1e-4);
            t += h;
            DLRA_CUDA(cudaMemcpyAsync(X, unew, N * sizeof(double), cudaMemcpyDeviceToDevice, cx.stream));
            std::swap(ks[0], ks[6]);   // FSAL
            st.naccept++;
            h = h / q;
            st.dt_next = h;
        } else {
            const double q11 = pow(EEst, beta1);
            const double q = std::min(1.0 / qmin, q11 / gamma);
            h = h / q;
            st.nreject++;
        }
    }
    if (carry) keep_stage(ks[0]);
}

// ------------------------------------------------------------------------------------------------------
// the three projected flows
// ------------------------------------------------------------------------------------------------------
// side flow: X' = Op·X + X·Msm + Fc·Fsm + c·had(D1X|X, D2X|X; T)   — K-flow (side = n) and L-flow (side = m)
struct SideFlow : FlowRhs {
    dlra_engine* e; DeWork* w;
    int64_t rows; int r;
    const dlra_operator* op;      // A (K-flow) or B (L-flow); may be NONE
    const double* Msm;            // r x r, ld r : X·Msm
    const double* Fc; int64_t ldf; int q; const double* Fsm;   // forcing: Fc (rows x q) · Fsm (q x r, ld q)
    double c_had; const double* T; // r^3 tensor or null
    const dlra_operator* d1; const dlra_operator* d2;           // K-flow: applied to X; L-flow: null (operands are X itself)
    double* s1; double* s2;        // rows x r scratch for D1·X, D2·X
    // two-sided terms Σ_k Op_k·X·Tsm_k: Op_k = A_k (K-flow) or B_k (L-flow), Tsm_k = VᵀB_kᵀV or UᵀA_kᵀU (r x r, ld r, packed)
    const std::vector<std::pair<dlra_operator, dlra_operator>>* terms = nullptr;
    bool side_is_n = true; const double* Tsm = nullptr; double* s3 = nullptr;
    void eval(const double* X, double* out, double) override {
        Ctx& cx = e->cx;
        if (op && op->kind != DLRA_OP_NONE) apply_op(cx, *op, rows, r, X, rows, out, rows, 1.0, 0.0);
        else fill_mat(cx, rows, r, out, rows, 0.0, 0.0);
        gemm_nn(cx, rows, r, r, X, rows, nullptr, 0, Msm, r, false, out, rows, 1.0, 1.0);
        if (q > 0) gemm_nn(cx, rows, q, r, Fc, ldf, nullptr, 0, Fsm, q, false, out, rows, 1.0, 1.0);
        if (c_had != 0.0) {
            const double* P = X; const double* Qm = X;
            if (d1) { apply_op(cx, *d1, rows, r, X, rows, s1, rows, 1.0, 0.0); P = s1; }
            if (d2) { apply_op(cx, *d2, rows, r, X, rows, s2, rows, 1.0, 0.0); Qm = s2; }
            if (r <= 32 && rows >= 2048) {
                hadamard_rows32_kernel<<<(unsigned)cdiv(rows, 64), 64, 0, cx.stream>>>(rows, r, P, rows, Qm, rows, T, out, rows, c_had);
            } else {
                dim3 grid((unsigned)cdiv(rows, 128), (unsigned)r);
                hadamard_rows_kernel<<<grid, 128, 0, cx.stream>>>(rows, r, P, rows, Qm, rows, T, out, rows, c_had);
            }
            cx.launches++;
            DLRA_CUDA(cudaGetLastError());
        }
        if (terms)
            for (size_t k = 0; k < terms->size(); ++k) {
                const dlra_operator& opk = side_is_n ? (*terms)[k].first : (*terms)[k].second;
                gemm_nn(cx, rows, r, r, X, rows, nullptr, 0, Tsm + (int64_t)k * r * r, r, false, s3, rows, 1.0, 0.0);   // X·Tsm_k
                apply_op(cx, opk, rows, r, s3, rows, out, rows, 1.0, 1.0);                                             // += Op_k·(X·Tsm_k)
            }
    }
};

inline void de_K_flow(dlra_engine* e, double* K, int r, const double* V, double t, double dt) {
    DeWork* w = de_work(e);
    Ctx& cx = e->cx;
    const RhsCfg& R = e->rhs;
    const int64_t n = e->n, m = e->m;
    DLRA_REQUIRE(R.set, "no right-hand side installed (dlra_rhs_set)");
    w->small.ensure(4 * (int64_t)e->W * e->W + (int64_t)R.q * e->W + 64, cx.stream);
    double* Bv = w->small.p;                                   // r x r : VᵀBᵀV
    double* Hv = Bv + (int64_t)e->W * e->W;                    // q x r : HᵀV
    w->mbuf[0].ensure(m * (int64_t)r, cx.stream);
    e->gws.ensure(std::max(gemm_tn_ws(cx, m, r, r), gemm_tn_ws(cx, m, std::max(R.q, 1), r)), cx.stream);
    if (R.B.kind != DLRA_OP_NONE) {
        apply_op(cx, R.B, m, r, V, m, w->mbuf[0].p, m, 1.0, 0.0);                              // B·V
        gemm_tn(cx, m, r, r, w->mbuf[0].p, m, nullptr, 0, V, m, Bv, r, 1.0, 0.0, e->gws.p);    // (BV)ᵀV = VᵀBᵀV
    } else {
        fill_mat(cx, r, r, Bv, r, 0.0, 0.0);
    }
    if (R.q > 0) gemm_tn(cx, m, R.q, r, R.H, R.ldh, nullptr, 0, V, m, Hv, R.q, 1.0, 0.0, e->gws.p);
    double* T = nullptr;
    if (R.c_had != 0.0) {
        w->tens[0].ensure((int64_t)r * r * r, cx.stream);
        tensor3(e, w, m, r, V, m, V, m, V, m, w->tens[0].p);
        T = w->tens[0].p;
        w->nbuf[1].ensure(n * (int64_t)r, cx.stream);
        w->nbuf[2].ensure(n * (int64_t)r, cx.stream);
    }
    SideFlow f;
    f.e = e; f.w = w; f.rows = n; f.r = r; f.op = &R.A; f.Msm = Bv; f.Fc = R.G; f.ldf = R.ldg; f.q = R.q; f.Fsm = Hv;
    f.c_had = R.c_had; f.T = T;
    f.d1 = (R.c_had != 0.0 && R.D1.kind != DLRA_OP_IDENTITY_SCALED) ? &R.D1 : nullptr;
    f.d2 = (R.c_had != 0.0 && R.D2.kind != DLRA_OP_IDENTITY_SCALED) ? &R.D2 : nullptr;
    if (R.c_had != 0.0) {
        // identity-scaled operands fold into the coefficient
        if (R.D1.kind == DLRA_OP_IDENTITY_SCALED) f.c_had *= R.D1.scale;
        if (R.D2.kind == DLRA_OP_IDENTITY_SCALED) f.c_had *= R.D2.scale;
    }
    f.s1 = w->nbuf[1].p; f.s2 = w->nbuf[2].p;
    if (!R.terms.empty()) {   // Tsm_k = VᵀB_kᵀV = (B_k·V)ᵀ·V
        w->tsm.ensure((int64_t)R.terms.size() * r * r, cx.stream);
        w->nbuf[3].ensure(n * (int64_t)r, cx.stream);
        for (size_t k = 0; k < R.terms.size(); ++k) {
            apply_op(cx, R.terms[k].second, m, r, V, m, w->mbuf[0].p, m, 1.0, 0.0);
            gemm_tn(cx, m, r, r, w->mbuf[0].p, m, nullptr, 0, V, m, w->tsm.p + (int64_t)k * r * r, r, 1.0, 0.0, e->gws.p);
        }
        f.terms = &R.terms; f.side_is_n = true; f.Tsm = w->tsm.p; f.s3 = w->nbuf[3].p;
    }
    // K lives in an engine buffer with ld == n (dense): integrate in place
    ode_advance(e, e->sub[DLRA_FLOW_K], f, n * (int64_t)r, K, t, dt);
}

inline void de_L_flow(dlra_engine* e, double* L, int r, const double* U, double t, double dt, FsalCarry* carry = nullptr) {
    DeWork* w = de_work(e);
    Ctx& cx = e->cx;
    const RhsCfg& R = e->rhs;
    const int64_t n = e->n, m = e->m;
    DLRA_REQUIRE(R.set, "no right-hand side installed (dlra_rhs_set)");
    w->small.ensure(4 * (int64_t)e->W * e->W + (int64_t)R.q * e->W + 64, cx.stream);
    double* Au = w->small.p + 2 * (int64_t)e->W * e->W;        // r x r : UᵀAᵀU
    double* Gu = Au + (int64_t)e->W * e->W;                    // q x r : GᵀU
    w->nbuf[0].ensure(n * (int64_t)r, cx.stream);
    e->gws.ensure(std::max(gemm_tn_ws(cx, n, r, r), gemm_tn_ws(cx, n, std::max(R.q, 1), r)), cx.stream);
    if (R.A.kind != DLRA_OP_NONE) {
        apply_op(cx, R.A, n, r, U, n, w->nbuf[0].p, n, 1.0, 0.0);                              // A·U
        gemm_tn(cx, n, r, r, w->nbuf[0].p, n, nullptr, 0, U, n, Au, r, 1.0, 0.0, e->gws.p);    // (AU)ᵀU = UᵀAᵀU
    } else {
        fill_mat(cx, r, r, Au, r, 0.0, 0.0);
    }
    if (R.q > 0) gemm_tn(cx, n, R.q, r, R.G, R.ldg, nullptr, 0, U, n, Gu, R.q, 1.0, 0.0, e->gws.p);
    double* T = nullptr;
    double chad = R.c_had;
    if (R.c_had != 0.0) {
        // T_U[a,b,c] = Σ_i (D1U)[i,a] (D2U)[i,b] U[i,c]
        w->nbuf[1].ensure(n * (int64_t)r, cx.stream);
        w->nbuf[2].ensure(n * (int64_t)r, cx.stream);
        apply_op(cx, R.D1, n, r, U, n, w->nbuf[1].p, n, 1.0, 0.0);
        apply_op(cx, R.D2, n, r, U, n, w->nbuf[2].p, n, 1.0, 0.0);
        w->tens[1].ensure((int64_t)r * r * r, cx.stream);
        tensor3(e, w, n, r, w->nbuf[1].p, n, w->nbuf[2].p, n, U, n, w->tens[1].p);
        T = w->tens[1].p;
    }
    SideFlow f;
    f.e = e; f.w = w; f.rows = m; f.r = r; f.op = &R.B; f.Msm = Au; f.Fc = R.H; f.ldf = R.ldh; f.q = R.q; f.Fsm = Gu;
    f.c_had = chad; f.T = T; f.d1 = nullptr; f.d2 = nullptr; f.s1 = nullptr; f.s2 = nullptr;
    if (!R.terms.empty()) {   // Tsm_k = UᵀA_kᵀU = (A_k·U)ᵀ·U
        w->tsm.ensure((int64_t)R.terms.size() * r * r, cx.stream);
        w->mbuf[1].ensure(m * (int64_t)r, cx.stream);
        for (size_t k = 0; k < R.terms.size(); ++k) {
            apply_op(cx, R.terms[k].first, n, r, U, n, w->nbuf[0].p, n, 1.0, 0.0);
            gemm_tn(cx, n, r, r, w->nbuf[0].p, n, nullptr, 0, U, n, w->tsm.p + (int64_t)k * r * r, r, 1.0, 0.0, e->gws.p);
        }
        f.terms = &R.terms; f.side_is_n = false; f.Tsm = w->tsm.p; f.s3 = w->mbuf[1].p;
    }
    ode_advance(e, e->sub[DLRA_FLOW_L], f, m * (int64_t)r, L, t, dt, carry);
}

// core flow: S' = sign·( Auu·S + S·Bvv + GH + c·contract(TU, S, S, TV) ),  S: p x q dense
struct CoreFlow : FlowRhs {
    dlra_engine* e; int p, q; double sign;
    const double* Auu; const double* Bvv; const double* GH; bool has_gh;
    double c_had; const double* TU; const double* TV; double* Y1; double* Y2;
    // two-sided terms: Σ_k (UᵀA_kU)·S·(VᵀB_kᵀV); Tuu: nterms x (p x p, ld p), Tvv: nterms x (q x q, ld q), Ttmp: p x q
    int nterms = 0; const double* Tuu = nullptr; const double* Tvv = nullptr; double* Ttmp = nullptr;
    void eval(const double* S, double* out, double) override {
        Ctx& cx = e->cx;
        small_gemm(cx, p, q, p, Auu, p, false, S, p, false, out, p, sign, 0.0);
        small_gemm(cx, p, q, q, S, p, false, Bvv, q, false, out, p, sign, 1.0);
        if (has_gh) copy_mat(cx, p, q, GH, p, false, out, p, sign, 1.0);
        if (c_had != 0.0) {
            // Y1[a', (b,d)] = Σ_a S[a',a] TV[a,(b,d)]            (p x q^2)
            small_gemm(cx, p, q * q, q, S, p, false, TV, q, false, Y1, p, 1.0, 0.0);
            // Y2[(a'), b', d] = Σ_b S[b',b] Y1[a', b, d]: for each d: Y2_d (p x p) = Y1_d (p x q) · Sᵀ (q x p)
            small_gemm(cx, p, p, q, Y1, p, false, S, p, true, Y2, p, 1.0, 0.0, q, (int64_t)p * q, 0, (int64_t)p * p);   // one batched launch
            // out[c,d] += sign·c_had · Σ_{a',b'} TU[(a',b'), c] · Y2[(a',b'), d]     (TUᵀ·Y2 with p^2 rows)
            small_gemm(cx, p, q, p * p, TU, p * p, true, Y2, p * p, false, out, p, sign * c_had, 1.0);
        }
        for (int k = 0; k < nterms; ++k) {
            small_gemm(cx, p, q, p, Tuu + (int64_t)k * p * p, p, false, S, p, false, Ttmp, p, 1.0, 0.0);
            small_gemm(cx, p, q, q, Ttmp, p, false, Tvv + (int64_t)k * q * q, q, false, out, p, sign, 1.0);
        }
    }
};

inline void de_S_flow(dlra_engine* e, double* S, int p, int q, const double* U, const double* V, double sign, double t, double dt) {
    DeWork* w = de_work(e);
    Ctx& cx = e->cx;
    const RhsCfg& R = e->rhs;
    const int64_t n = e->n, m = e->m, W = e->W;
    DLRA_REQUIRE(R.set, "no right-hand side installed (dlra_rhs_set)");
    const int64_t sm_need = 8 * W * W + 2 * (int64_t)R.q * W + 64;
    w->small.ensure(sm_need, cx.stream);
    double* Auu = w->small.p + 4 * W * W;     // p x p (ld p)
    double* Bvv = Auu + W * W;                // q x q (ld q)
    double* GH = Bvv + W * W;                 // p x q (ld p)
    double* Sd = GH + W * W;                  // dense copy of S (p x q, ld p)
    double* Gu = Sd + W * W;                  // qf x p
    double* Hv = Gu + (int64_t)R.q * W;       // qf x q
    w->nbuf[0].ensure(n * (int64_t)p, cx.stream);
    w->mbuf[0].ensure(m * (int64_t)q, cx.stream);
    e->gws.ensure(std::max(std::max(gemm_tn_ws(cx, n, p, p), gemm_tn_ws(cx, m, q, q)),
                           std::max(gemm_tn_ws(cx, n, std::max(R.q, 1), p), gemm_tn_ws(cx, m, std::max(R.q, 1), q))), cx.stream);
    if (R.A.kind != DLRA_OP_NONE) {
        apply_op(cx, R.A, n, p, U, n, w->nbuf[0].p, n, 1.0, 0.0);
        gemm_tn(cx, n, p, p, U, n, nullptr, 0, w->nbuf[0].p, n, Auu, p, 1.0, 0.0, e->gws.p);       // Uᵀ(AU)
    } else fill_mat(cx, p, p, Auu, p, 0.0, 0.0);
    if (R.B.kind != DLRA_OP_NONE) {
        apply_op(cx, R.B, m, q, V, m, w->mbuf[0].p, m, 1.0, 0.0);
        gemm_tn(cx, m, q, q, w->mbuf[0].p, m, nullptr, 0, V, m, Bvv, q, 1.0, 0.0, e->gws.p);       // (BV)ᵀV = VᵀBᵀV
    } else fill_mat(cx, q, q, Bvv, q, 0.0, 0.0);
    const bool has_gh = R.q > 0;
    if (has_gh) {
        gemm_tn(cx, n, R.q, p, R.G, R.ldg, nullptr, 0, U, n, Gu, R.q, 1.0, 0.0, e->gws.p);         // GᵀU (qf x p)
        gemm_tn(cx, m, R.q, q, R.H, R.ldh, nullptr, 0, V, m, Hv, R.q, 1.0, 0.0, e->gws.p);         // HᵀV (qf x q)
        small_gemm(cx, p, q, R.q, Gu, R.q, true, Hv, R.q, false, GH, p, 1.0, 0.0);                 // (UᵀG)(HᵀV)
    }
    CoreFlow f;
    f.e = e; f.p = p; f.q = q; f.sign = sign; f.Auu = Auu; f.Bvv = Bvv; f.GH = GH; f.has_gh = has_gh;
    f.c_had = R.c_had; f.TU = nullptr; f.TV = nullptr; f.Y1 = nullptr; f.Y2 = nullptr;
    if (R.c_had != 0.0) {
        w->nbuf[1].ensure(n * (int64_t)p, cx.stream);
        w->nbuf[2].ensure(n * (int64_t)p, cx.stream);
        apply_op(cx, R.D1, n, p, U, n, w->nbuf[1].p, n, 1.0, 0.0);
        apply_op(cx, R.D2, n, p, U, n, w->nbuf[2].p, n, 1.0, 0.0);
        w->tens[1].ensure((int64_t)p * p * p, cx.stream);
        tensor3(e, w, n, p, w->nbuf[1].p, n, w->nbuf[2].p, n, U, n, w->tens[1].p);
        w->tens[0].ensure((int64_t)q * q * q, cx.stream);
        tensor3(e, w, m, q, V, m, V, m, V, m, w->tens[0].p);
        w->tens[2].ensure((int64_t)p * q * q + (int64_t)p * p * q, cx.stream);
        f.TU = w->tens[1].p; f.TV = w->tens[0].p; f.Y1 = w->tens[2].p; f.Y2 = w->tens[2].p + (int64_t)p * q * q;
    }
    if (!R.terms.empty()) {
        const int64_t nt = (int64_t)R.terms.size();
        w->tsm.ensure(nt * ((int64_t)p * p + (int64_t)q * q) + (int64_t)p * q, cx.stream);
        double* Tuu = w->tsm.p; double* Tvv = Tuu + nt * p * p; double* Ttmp = Tvv + nt * q * q;
        for (int64_t k = 0; k < nt; ++k) {
            apply_op(cx, R.terms[k].first, n, p, U, n, w->nbuf[0].p, n, 1.0, 0.0);
            gemm_tn(cx, n, p, p, U, n, nullptr, 0, w->nbuf[0].p, n, Tuu + k * p * p, p, 1.0, 0.0, e->gws.p);          // Uᵀ(A_k U)
            apply_op(cx, R.terms[k].second, m, q, V, m, w->mbuf[0].p, m, 1.0, 0.0);
            gemm_tn(cx, m, q, q, w->mbuf[0].p, m, nullptr, 0, V, m, Tvv + k * q * q, q, 1.0, 0.0, e->gws.p);          // (B_k V)ᵀV
        }
        f.nterms = (int)nt; f.Tuu = Tuu; f.Tvv = Tvv; f.Ttmp = Ttmp;
    }
    copy_mat(cx, p, q, S, W, false, Sd, p);
    ode_advance(e, e->sub[DLRA_FLOW_S], f, (int64_t)p * q, Sd, t, dt);
    copy_mat(cx, p, q, Sd, p, false, S, W);
}

}  // namespace dlra
