// Engine state behind a dlra_handle: device-resident factors, workspaces sized once for rmax
// (no alg_recache-style reallocation, cf. rank_adaptive_unconventional.jl:133-169), the data feed and streams.
#pragma once
#include "common.cuh"
#include "comm.cuh"
#include "small_ops.cuh"
#include "tsqr.cuh"
#include "jacobi.cuh"
#include "../../include/dlra.h"
#include <vector>
#include <algorithm>

namespace dlra {

struct DevBuf {
    double* p = nullptr;
    int64_t n = 0;  // doubles
    void ensure(int64_t want, cudaStream_t s) {
        if (want <= n) return;
        if (p) { DLRA_CUDA(cudaStreamSynchronize(s)); DLRA_CUDA(cudaFree(p)); p = nullptr; n = 0; }
        cudaError_t e = cudaMalloc(&p, (size_t)want * sizeof(double));
        if (e != cudaSuccess) { p = nullptr; throw CudaError(5, std::string("cudaMalloc of ") + std::to_string(want * 8) + " bytes failed: " + cudaGetErrorString(e)); }
        n = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// ΔA = A − Aprev (Aprev == nullptr: A is the increment itself, or the full snapshot for the greedy step)
struct Delta {
    const double* A = nullptr;
    int64_t lda = 0;
    const double* Aprev = nullptr;
    int64_t ldap = 0;
};

struct SubStepperCfg {
    int ode = DLRA_ODE_TSIT5;
    int nsub = 1;
    double abstol = 1e-6, reltol = 1e-3;
    int64_t maxiters = 100000;   // OrdinaryDiffEq's default `maxiters` (sub-steps incl. rejected ones per `step!(I, dt, true)`)
    // adaptive controller state carried across outer steps (mirrors oracle.SubStepper)
    double dt_next = -1.0;
    double qold = 1e-4;
    int64_t nfev = 0, naccept = 0, nreject = 0;
};

// first stage kept across outer steps by an integrator that is never `set_u!`-ed (hybrid Z-flow, greedy_integrator.jl:72-76)
struct FsalCarry {
    DevBuf k;
    bool valid = false;
    int64_t N = 0;
};

struct RhsCfg {
    bool set = false;
    dlra_operator A{}, B{}, D1{}, D2{};
    const double* G = nullptr; int64_t ldg = 0;
    const double* H = nullptr; int64_t ldh = 0;
    int q = 0;
    double c_had = 0.0;
    // two-sided terms  Σ_k A_k·X·B_kᵀ  (dlra_rhs_add_term)
    std::vector<std::pair<dlra_operator, dlra_operator>> terms;
};

}  // namespace dlra

struct dlra_engine {
    int device = 0;
    int64_t n = 0, m = 0;       // local rows, columns
    int r = 0, rmax = 0, flags = 0;
    int W = 0;                  // widest factor block (rmax, or 2*rmax when rank adaptive)
    dlra::Ctx cx;
    dlra::Ctx ax;               // auxiliary stream: the replicated m-side chain (QR(L), N) overlaps the n-side chain (QR(K), M)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_ext = nullptr;
    cudaEvent_t ev_rnew = nullptr;      // rank-adaptive step: the new rank has landed in r_new_host
    cudaEvent_t ev_kqr = nullptr, ev_join2 = nullptr;   // pipelined BUG step: K-side QR done (main) / M = U1'U0 done (auxiliary)
    dlra::DevBuf UC;                    // third n x W factor buffer of the pipelined BUG step (allocated on first use)
    // DLRA_AUG_BASIS_FIRST takes the leading panel [U0], [V0] of the augmented bases as orthonormal: true for factors produced by a
    // step of this engine, unknown for factors handed in through dlra_set_factors; every AUG_REORTHO_EVERY-th step factors the whole
    // augmented basis anyway so that the orthogonality defect cannot accumulate over long runs
    bool basis_trusted = false;
    int64_t aug_steps = 0;
    static constexpr int AUG_REORTHO_EVERY = 256;
    dlra::Comm comm;            // row-shard communicator
    dlra::Comm self;            // nranks = 1: for replicated (m-side) factorizations
    cudaStream_t copy_stream = nullptr;
    std::string err;

    // factors (ld: U -> n, V -> m, small -> W)
    double *U = nullptr, *UB = nullptr, *V = nullptr, *VB = nullptr, *S = nullptr;
    // small matrices, each W x W, ld = W
    double *M = nullptr, *N = nullptr, *Sh = nullptr, *T1 = nullptr, *T2 = nullptr, *Rm = nullptr, *Pm = nullptr, *Qm = nullptr, *sig = nullptr, *stg = nullptr;
    double* small_block = nullptr;
    int* r_new_dev = nullptr;
    int* r_new_host = nullptr;  // pinned, mapped
    double* scal_dev = nullptr; // 8 doubles of device scalars
    dlra::DevBuf gws, tws, wtmp, jws, nscr, mscr, part;  // grow-on-demand scratch
    dlra::DevBuf gws2, tws2, wtmp2;                       // scratch of the auxiliary stream
    dlra::DevBuf isvd;                                    // temporaries of dlra_truncated_svd
    dlra::DevBuf bstage;                                  // staged small operand of tall_gemm_tma

    // data feed
    const double* prev = nullptr; int64_t ldprev = 0;
    const double* cur = nullptr; int64_t ldcur = 0; int cur_kind = 0; bool have_cur = false;
    const double* nxt = nullptr; int64_t ldnxt = 0; int nxt_kind = 0; bool have_nxt = false;   // one-snapshot lookahead
    static constexpr int NOWN = 4;
    double* own[NOWN] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t own_free[NOWN] = {nullptr, nullptr, nullptr, nullptr};   // recorded on the compute stream when a step stops reading own[i]
    cudaEvent_t own_ready[NOWN] = {nullptr, nullptr, nullptr, nullptr};  // recorded on the copy stream when the H2D copy landed
    int own_next = 0; int cur_own = -1; int prev_own = -1; int nxt_own = -1;
    // software pipelining of the BUG step (pass_tri.cuh): ΔA·V0 already sits in UB and the per-CTA partials of ΔAᵀ·U0 in `part`
    bool kl_ready = false; int kl_nparts = 0; int64_t kl_ldlp = 0; int kl_rank = 0;
    bool kl_lsum_done = false;   // ... and the fixed-order sum of those partials already sits in VB (formed beside the core update)
    cudaEvent_t ev_pass = nullptr, ev_lsum = nullptr;
    bool lsum_pending = false;   // the early L sum may still be running on the auxiliary stream (it writes VB): see settle_aux

    // asynchronous factor snapshots (dlra_save_factors_async): two device staging slots, D2H on the copy stream
    dlra::DevBuf save_stage[2];
    cudaEvent_t save_staged[2] = {nullptr, nullptr};   // compute stream: staging copy of slot i complete
    cudaEvent_t save_landed[2] = {nullptr, nullptr};   // copy stream: D2H of slot i complete (the slot may be overwritten)
    int save_next = 0;

    // DE problems
    dlra::RhsCfg rhs;
    dlra::SubStepperCfg sub[3];
    dlra::FsalCarry zcarry;     // hybrid Z-flow (uses sub[DLRA_FLOW_L])

    // profiling of the dominant contraction kernels
    bool time_passes = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pass_events;
    std::vector<int> pass_event_kind;
    int64_t pass_launches = 0;
    double pass_bytes = 0.0, pass_ms = 0.0;
    int64_t kind_launches[4] = {0, 0, 0, 0};
    double kind_ms[4] = {0, 0, 0, 0}, kind_bytes[4] = {0, 0, 0, 0}, kind_flops[4] = {0, 0, 0, 0};
    cudaEvent_t user_events[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // step progress (dlra_progress): one event per step in a ring, polled without blocking
    static constexpr int NPROG = 64;
    cudaEvent_t prog_ev[NPROG] = {};
    int64_t steps_enqueued = 0, steps_completed = 0;
    // DLRA_PHASES=1: CUDA-event marks at the phase boundaries of a step on the main stream; the per-phase means are printed to
    // stderr by dlra_stats (the time between two marks is attributed to the later one)
    bool phase_timing = false;
    std::vector<std::pair<const char*, cudaEvent_t>> phase_marks;
};

namespace dlra {
inline void phase_mark(dlra_engine* e, const char* name) {
    if (!e->phase_timing) return;
    cudaEvent_t ev;
    DLRA_CUDA(cudaEventCreate(&ev));
    DLRA_CUDA(cudaEventRecord(ev, e->cx.stream));
    e->phase_marks.emplace_back(name, ev);
}
inline void pass_timer_begin(dlra_engine* e, double bytes, int kind = 1, double flops = 0.0) {
    e->pass_launches++;
    e->pass_bytes += bytes;
    e->kind_launches[kind]++;
    e->kind_bytes[kind] += bytes;
    e->kind_flops[kind] += flops;
    if (!e->time_passes) return;
    e->pass_event_kind.push_back(kind);
    cudaEvent_t a, b;
    DLRA_CUDA(cudaEventCreate(&a));
    DLRA_CUDA(cudaEventCreate(&b));
    DLRA_CUDA(cudaEventRecord(a, e->cx.stream));
    e->pass_events.emplace_back(a, b);
}
inline void pass_timer_end(dlra_engine* e) {
    if (!e->time_passes) return;
    DLRA_CUDA(cudaEventRecord(e->pass_events.back().second, e->cx.stream));
}

}  // namespace dlra
