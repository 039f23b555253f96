/*
 * dlra.h — C ABI of libdlra.so, the B200-native engine for the per-step DLRA hot path of
 * FHoltorf/LowRankIntegrators.jl.  The reference has no FFI: its seam is Julia
 * multiple dispatch on  alg_cache(prob, alg, u, dt) / step!(integrator, alg, dt)
 * (src/integrators/projector_splitting.jl:43,87,191-211; unconventional.jl:39,86,159-164;
 * rank_adaptive_unconventional.jl:47,106,171-180).  Each entry point below names the reference
 * code it replaces; julia/DLRAB200.jl binds them with `ccall`, and
 * lowrankintegrators.jl_b200/_lib.py binds the same symbols with ctypes.
 *
 * Conventions: all matrices are fp64, column-major (Julia layout) with an explicit leading
 * dimension in elements; pointers are DEVICE pointers unless the function name ends in `_host`;
 * every call returns 0 on success or a DLRA_E* code, with the text in dlra_last_error();
 * work is enqueued on the engine's own CUDA stream and is asynchronous unless stated;
 * one caller thread per handle (the reference is single-task), handles are independent.
 * One handle drives ONE GPU; multi-GPU runs use one process (handle) per GPU with the matrix
 * rows (n) sharded in contiguous blocks, joined by dlra_comm_init (NCCL over NVLink).
 */
#ifndef DLRA_H_
#define DLRA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dlra_engine* dlra_handle;

enum {
    DLRA_OK = 0,
    DLRA_EINVAL = 1,     /* bad argument (the reference would throw MethodError / AssertionError) */
    DLRA_ECUDA = 2,      /* CUDA runtime / driver failure */
    DLRA_ENCCL = 3,      /* NCCL failure or NCCL not loadable */
    DLRA_ESTATE = 4,     /* call sequence error (e.g. step without pushed data / without rhs) */
    DLRA_ENOMEM = 5,
    DLRA_EUNSUPPORTED = 6,
    DLRA_EMAXITERS = 7   /* adaptive sub-stepper hit its maxiters (OrdinaryDiffEq would return retcode MaxIters); the factors and
                          * the controller state are those before the failed step */
};

/* dlra_create flags */
enum {
    DLRA_RANK_ADAPTIVE = 1,  /* size workspaces for the augmented 2r bases of the rank-adaptive integrator */
    DLRA_FORCE_GENERIC = 2,  /* debugging: use the generic (non-TMA, non-DMMA) contraction kernels only */
    DLRA_AUG_BASIS_FIRST = 4 /* rank-adaptive step: factor the augmented bases as [U0 | K], [V0 | L] instead of [K | U0], [L | V0]
                              * (same span, hence the same U·S·Vᵀ and ranks; the leading panel is already orthonormal and needs no
                              * TSQR).  Opt-in: not validated on hardware in round 1. */
};

/* KSL orders: PrimalLieTrotter / DualLieTrotter / Strang (projector_splitting.jl:1-3) */
enum { DLRA_KSL_PRIMAL = 0, DLRA_KSL_DUAL = 1, DLRA_KSL_STRANG = 2 };

/* kinds of pushed data (src/integrators/data_integrator.jl:22-28 + the ΔA formation at
 * projector_splitting.jl:119-121 and twins) */
enum {
    DLRA_DATA_SNAPSHOT = 0, /* A(t+dt); the engine forms ΔA = A(t+dt) − A(t) on the fly and then yprev ← ycurr */
    DLRA_DATA_DELTA = 1     /* a pre-differenced increment ΔA */
};

/* sub-flow selectors and explicit RK sub-steppers standing in for K_alg/S_alg/L_alg
 * (projector_splitting.jl:36-38, unconventional.jl:16-18, rank_adaptive_unconventional.jl:18-20) */
enum { DLRA_FLOW_K = 0, DLRA_FLOW_S = 1, DLRA_FLOW_L = 2 };
enum { DLRA_ODE_EULER = 0, DLRA_ODE_RK4 = 1, DLRA_ODE_TSIT5_FIXED = 2, DLRA_ODE_TSIT5 = 3 };

/* operator storage for the device-evaluable right-hand sides */
enum { DLRA_OP_NONE = 0, DLRA_OP_DENSE = 1, DLRA_OP_CSR = 2, DLRA_OP_IDENTITY_SCALED = 3 };

typedef struct dlra_operator {
    int kind;              /* DLRA_OP_* */
    int64_t rows, cols;
    const double* dense;   /* DENSE: rows x cols column-major, ld */
    int64_t ld;
    const int64_t* rowptr; /* CSR (0-based, rows+1 entries) */
    const int32_t* colind;
    const double* values;
    double scale;          /* IDENTITY_SCALED: scale * I ; otherwise multiplies the operator */
} dlra_operator;

/* ---- lifetime -------------------------------------------------------------------------------- */

/* Replaces alg_cache(::MatrixDataProblem|::MatrixDEProblem, alg, u, dt) buffer allocation
 * (projector_splitting.jl:43-105, unconventional.jl:39-107, rank_adaptive_unconventional.jl:47-131,
 * and alg_recache :133-169 — workspaces are sized once for rmax, never reallocated per rank change).
 * n_local: rows of this process' shard (== n on one GPU); m: columns; r0: initial rank; rmax: rank cap. */
int dlra_create(int device, int64_t n_local, int64_t m, int r0, int rmax, int flags, dlra_handle* out);
int dlra_destroy(dlra_handle h);
const char* dlra_last_error(dlra_handle h); /* h may be NULL: last error of a failed dlra_create */
const char* dlra_version(void);

/* ---- multi-GPU (no counterpart in the single-process reference) ---------------- */
int dlra_nccl_unique_id(void* id128);                       /* 128-byte ncclUniqueId, made on rank 0 */
int dlra_comm_init(dlra_handle h, int nranks, int rank, const void* id128);
/* Peer-to-peer transport (preferred on one NVSwitch box): every rank exports the CUDA-IPC handle (64 bytes) of its
 * exchange region, the host runtime all-gathers the handles, every rank imports all of them.  Afterwards the
 * collectives of the step run as single kernels that read the peers' contributions over NVLink (no NCCL calls). */
int dlra_p2p_export(dlra_handle h, void* ipc_handle64);
int dlra_p2p_import(dlra_handle h, int nranks, int rank, const void* ipc_handles /* nranks x 64 bytes */);

/* ---- factors: SVDLikeRepresentation(U,S,V) (LowRankArithmetic; README.md:85) ----------------- */
/* deep copy in, as init() does with deepcopy(prob.u0) (projector_splitting.jl:110) */
int dlra_set_factors_host(dlra_handle h, const double* U, int64_t ldu, const double* S, int64_t lds,
                          const double* V, int64_t ldv, int r);
int dlra_set_factors(dlra_handle h, const double* U, int64_t ldu, const double* S, int64_t lds,
                     const double* V, int64_t ldv, int r);
/* update_sol! (primitives.jl:82-90): deep copy out; synchronises the engine stream */
int dlra_get_factors_host(dlra_handle h, double* U, int64_t ldu, double* S, int64_t lds, double* V,
                          int64_t ldv, int* r);
int dlra_get_factors(dlra_handle h, double* U, int64_t ldu, double* S, int64_t lds, double* V,
                     int64_t ldv, int* r);
/* Asynchronous `update_sol!` (primitives.jl:82-90: `deepcopy(integrator.u)` after every step) / save-at: snapshot the
 * current factors into HOST buffers (pinned for a truly asynchronous copy) without stalling the step stream — staged on the
 * device, moved by the copy stream while later steps run.  *r receives the rank immediately; the buffers are valid after
 * dlra_save_wait (or dlra_sync + dlra_save_wait).  At most two snapshots are in flight; a third waits on the device. */
int dlra_save_factors_async(dlra_handle h, double* U_host, int64_t ldu, double* S_host, int64_t lds, double* V_host,
                            int64_t ldv, int* r);
int dlra_save_wait(dlra_handle h);
int dlra_get_rank(dlra_handle h, int* r);
/* truncated_svd(A, r) / truncated_svd(A; tol) (LowRankArithmetic; call sites test/data_driven_approximation.jl:18,
 * examples/generic_matrix.jl:29) for matrices that only exist on the device: randomized subspace
 * iteration with the engine's own kernels (K-only / L-only passes, TSQR, Jacobi SVD) — A (n_local x m, device) is
 * streamed 2*(power_iters+1) times, no n x m SVD is formed.  r > 0 fixes the rank; r == 0 picks it from `tol` with the
 * truncate_to_tolerance rule among the first min(rmax, sketch) values.  sketch = r + oversample columns (<= 128).
 * The result (an approximation of the reference's exact LAPACK truncation, error within a few % of optimal for
 * decaying spectra) becomes the engine's factors.  Works row-sharded. */
int dlra_truncated_svd(dlra_handle h, const double* A, int64_t ld, int r, double tol, int oversample,
                       int power_iters, uint64_t seed);
/* borrow the engine's live factor buffers (valid until the next step; ld of U,V = n_local, m; ld of S = rmax cap) */
int dlra_factor_ptrs(dlra_handle h, const double** U, int64_t* ldu, const double** S, int64_t* lds,
                     const double** V, int64_t* ldv, int* r);

/* ---- data feed: MatrixDataProblem (primitives.jl:23-30; data_integrator.jl:22-28) ------------- */
/* yprev = y[1] | y(t0)   (projector_splitting.jl:91).  The *_host forms copy into engine-owned
 * device buffers through a dedicated copy stream (overlaps the running step); the device forms
 * BORROW the pointer: it must stay valid and unmodified until the step after the next push has run. */
int dlra_data_init(dlra_handle h, const double* A0, int64_t ld);
int dlra_data_init_host(dlra_handle h, const double* A0, int64_t ld);
/* update_data!(ycurr, y, t, dt) for the NEXT step; kind = DLRA_DATA_SNAPSHOT | DLRA_DATA_DELTA.
 * A SECOND push before the step is a one-snapshot lookahead (the data of the step after next): with it dlra_step_bug
 * forms the next step's K/L contractions in the same sweep as this step's core pass (each snapshot is then read three
 * times instead of four); results are identical.  The lookahead snapshot becomes the pushed data of the next step. */
int dlra_data_push(dlra_handle h, const double* A, int64_t ld, int kind);
int dlra_data_push_host(dlra_handle h, const double* A, int64_t ld, int kind);

/* ---- right-hand sides: MatrixDEProblem (primitives.jl:13-17) in device-evaluable form ---------- */
/* F(X,t) = A·X + X·Bᵀ + G·Hᵀ + c_had·(D1·X).*(D2·X)      (all terms optional)
 * covers examples/generic_matrix.jl-style linear problems (W1·X + X + X·W2), Lyapunov-type
 * A·X + X·Bᵀ (+ low-rank forcing) and the Burgers UQ right-hand side Δρ − (∇ρ).*ρ
 * (test/data_agnostic_approximation.jl:31-33).  Replaces the default K_rhs/L_rhs/S_rhs closures
 * (projector_splitting.jl:52-80, unconventional.jl:51-79, rank_adaptive_unconventional.jl:59-86).
 * Operators are n x n (A, D1, D2) and m x m (B); G is n x q, H is m x q (device, column-major).
 * With row sharding (dlra_comm_init) A, D1, D2, G take this rank's row block of the global operator
 * (rows = n_local, cols = n_global).  Pointers are borrowed for the life of the handle. */
int dlra_rhs_set(dlra_handle h, const dlra_operator* A, const dlra_operator* B, const double* G,
                 int64_t ldg, const double* H, int64_t ldh, int q, const dlra_operator* D1,
                 const dlra_operator* D2, double c_had);
/* Adds one two-sided term A_k·X·B_kᵀ to the installed right-hand side: F(X) += A_k·X·B_kᵀ (A_k: n x n, B_k: m x m, dense,
 * CSR or scale·I).  Covers generators of the chemical-master-equation type, examples/markov_chain.jl:64-66
 * (Σ_r A_r .* (S_r·P·T_r) − Asum .* P: the rank-one Hadamard weights are diagonal scalings folded into the sparse
 * operators).  dlra_rhs_set clears the list.  NOT VALIDATED ON HARDWARE in round 1 (tests gated by DLRA_UNVALIDATED=1). */
int dlra_rhs_add_term(dlra_handle h, const dlra_operator* A, const dlra_operator* B);
/* K_alg / S_alg / L_alg and their tolerances (Tsit5 defaults abstol=1e-6, reltol=1e-3) */
int dlra_set_substepper(dlra_handle h, int flow, int ode, int nsub, double abstol, double reltol);
/* `maxiters` of the adaptive sub-stepper of one flow (OrdinaryDiffEq default 100000, the value the reference's
 * `init(ODEProblem, Tsit5(); save_everystep=false)` integrators run with, unconventional.jl:59): accepted + rejected
 * sub-steps allowed inside one outer step before the step fails with DLRA_EMAXITERS. */
int dlra_set_substepper_maxiters(dlra_handle h, int flow, int64_t maxiters);

/* ---- steps: one call == one reference step!(integrator, alg, dt) minus the t/iter bookkeeping -- */
/* primal_LT_step! / dual_LT_step! / Strang (projector_splitting.jl:117-211).  For data problems
 * the pushed ΔA is consumed; Strang on data needs two pushes, so call PRIMAL and DUAL with dt/2. */
int dlra_step_ksl(dlra_handle h, int order, double t, double dt);
/* unconventional_step! (unconventional.jl:121-164) */
int dlra_step_bug(dlra_handle h, double t, double dt);
/* rankadaptive_unconventional_step! + alg_recache (rank_adaptive_unconventional.jl:133-233).
 * Synchronises (the new rank decides the next step's shapes). */
int dlra_step_rabug(dlra_handle h, double t, double dt, double tol, int64_t rmax, int* r_new,
                    int* rank_changed);
/* greedy_step!(::SVDLikeRepresentation, ..., ::MatrixDataProblem) (greedy_integrator.jl:94-104):
 * re-projection on the full pushed snapshot X (must be pushed as DLRA_DATA_SNAPSHOT). */
int dlra_step_greedy(dlra_handle h, double t, double dt);
/* greedy_step!(::TwoFactorRepresentation, ...) (greedy_integrator.jl:72-92), u = U*Z'.  The engine keeps Z in the V slot
 * and S at identity (set the factors as (U, I, Z); dlra_get_factors returns (U, I, Z), dlra_reconstruct gives U*Z').
 *   DLRA_GREEDY_DATA   (MatrixDataProblem, :84-92):   Z = X'*U with the full pushed snapshot X
 *   DLRA_GREEDY_HYBRID (MatrixHybridProblem, :72-82): Z advanced over [t, t+dt] by dZ/dt = F(U*Z')'*U with the right-hand
 *       side installed by dlra_rhs_set (the FZ of test/data_informed_approximation.jl:75) and the sub-stepper configured
 *       for DLRA_FLOW_L; the pushed snapshot is X = y(t+dt)
 * then U = Q*P' of svd(X*Z) (the orthogonal polar factor; TSQR + small SVD here).
 * carry_fsal != 0 keeps the Z integrator's cached first stage across calls like the reference's ZIntegrator, which is
 * never set_u!-ed (the stage was evaluated with the previous basis); 0 re-evaluates it with the current U. */
enum { DLRA_GREEDY_DATA = 0, DLRA_GREEDY_HYBRID = 1 };
int dlra_step_greedy_two_factor(dlra_handle h, int mode, int carry_fsal, double t, double dt);

/* Blocks until everything enqueued on the engine's compute stream has run.  Steps are asynchronous: device-side failures surface
 * here (DLRA_ECUDA), including a timed-out inter-CTA wait of the one-launch TSQR (factors computed since then are invalid). */
int dlra_sync(dlra_handle h);
/* The engine runs on its own (non-blocking) stream.  Device buffers handed to it must be complete: either synchronise the
 * producing stream on the host, or call this with that stream (a cudaStream_t; NULL = the legacy default stream) — the
 * engine stream then waits for everything enqueued there so far (event record + cudaStreamWaitEvent, no host blocking). */
int dlra_wait_stream(dlra_handle h, void* producer_stream);
/* The engine's compute stream (a cudaStream_t), for the opposite direction: a caller whose allocator recycles device memory
 * by stream (Julia's CUDA.jl pool, PyTorch's caching allocator: `tensor.record_stream`) registers borrowed snapshots with this
 * stream so that a freed snapshot is not handed out again while a queued step still reads it (data is borrowed read-only and
 * the steps are asynchronous; the reference's `y` is never freed under the integrator, primitives.jl:23-30). */
int dlra_get_stream(dlra_handle h, void** stream);
/* Steps are asynchronous: *steps_enqueued counts the dlra_step_* calls made so far, *steps_completed those whose work has
 * finished on the device (polled without blocking).  wait_for > *steps_completed blocks until that many steps are done
 * (pass -1 never to block).  A borrowed snapshot that served step k as "current" is last read by step k+1: callers that
 * recycle snapshot memory release it once steps_completed >= k+1. */
int dlra_progress(dlra_handle h, int64_t* steps_enqueued, int64_t* steps_completed, int64_t wait_for);

/* ---- diagnostics ------------------------------------------------------------------------------ */
/* ‖U·S·Vᵀ − Yref‖_F / ‖Yref‖_F without materialising n x m on the host (Yref: n_local x m device);
 * with row sharding both norms are all-reduced.  Synchronises. */
int dlra_reconstruct_error(dlra_handle h, const double* Yref, int64_t ld, double* rel_fro);
/* normal_component (utils.jl:2-20): N = (I − U·Uᵀ)·dY·(I − Z·pinv(C, atol = tol)·Zᵀ) with Z = V·Sᵀ of the current factors
 * (Z itself for the two-factor convention S = I) and C = ZᵀZ unless a device r x r matrix C is given.  dY: n_local x m
 * device.  fro_norm (host, may be NULL) receives ‖N‖_F over all row shards; out (device n_local x m, may be NULL)
 * receives N.  Two streaming passes over dY (one fused K/L contraction, one rank-2r downdate); discards a pending
 * lookahead contraction of the pipelined BUG step.  Synchronises when fro_norm is requested. */
int dlra_normal_component(dlra_handle h, const double* dY, int64_t ld, const double* C, int64_t ldc, double tol, double* out,
                          int64_t ldo, double* fro_norm);
/* dense reconstruction Y = U·S·Vᵀ into a device buffer (Matrix(u), LowRankArithmetic) */
int dlra_reconstruct(dlra_handle h, double* Y, int64_t ld);
/* number of kernels this handle launched so far / device ms of the dominant contraction kernel
 * accumulated since the last reset (CUDA events on the engine stream) */
int dlra_stats(dlra_handle h, int64_t* kernel_launches, int64_t* pass_launches, double* pass_ms_total,
               double* pass_bytes_total, int reset);
int dlra_set_profiling(dlra_handle h, int time_passes);
/* per-kind breakdown of the timed contraction launches since the last reset:
 * index 0 = fused K+L pass, 1 = K-only pass (also the first half of the S pass), 2 = L-only pass,
 * 3 = software-pipelined BUG pass (core of step k + K/L of step k+1 in one sweep over three snapshots) */
int dlra_pass_breakdown(dlra_handle h, int64_t launches[4], double ms[4], double bytes[4], double flops[4]);
/* CUDA events on the ENGINE's stream (torch.cuda.Event only sees torch's streams): slots 0..7 */
int dlra_event_record(dlra_handle h, int slot);
int dlra_event_elapsed_ms(dlra_handle h, int slot_begin, int slot_end, double* ms); /* synchronises on slot_end */

#ifdef __cplusplus
}
#endif
#endif /* DLRA_H_ */
