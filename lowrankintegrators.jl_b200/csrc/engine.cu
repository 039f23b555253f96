// libdlra.so — C ABI (include/dlra.h) and the per-step orchestration of the KSL / BUG / rank-adaptive BUG
// integrators on one B200.  Step algebra follows SURVEY.md Appendix A, which restates
// src/integrators/projector_splitting.jl:117-189, unconventional.jl:121-157 and
// rank_adaptive_unconventional.jl:182-233 of the reference.  No CPU fallback exists: every entry point
// needs a CUDA device.
#include "engine.cuh"
#include "passes.cuh"
#include "de_flows.cuh"

using namespace dlra;

static void phase_mark_hook(void* eng, const char* name) { phase_mark((dlra_engine*)eng, name); }
static bool tall_gemm_hook(void* eng, int64_t n, int p, int q, const double* A, int64_t lda, const double* B, int64_t ldb, bool transB,
                           double* C, int64_t ldc, double alpha, double beta) {
    return tall_gemm_tma((dlra_engine*)eng, n, p, q, A, lda, B, ldb, transB, C, ldc, alpha, beta);
}

static thread_local std::string g_create_error;

#define DLRA_API_BEGIN(h)                                                     \
    if (!(h)) { g_create_error = "null handle"; return DLRA_EINVAL; }         \
    try {                                                                     \
        DLRA_CUDA(cudaSetDevice((h)->device));

#define DLRA_API_END(h)                                                       \
    }                                                                         \
    catch (const CudaError& ex) { (h)->err = ex.what(); return ex.code; }     \
    catch (const std::exception& ex) { (h)->err = ex.what(); return DLRA_ECUDA; } \
    return DLRA_OK;

// ---------------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------------
extern "C" const char* dlra_version(void) { return "dlra-b200 0.1 (sm_100a)"; }

extern "C" const char* dlra_last_error(dlra_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int dlra_destroy(dlra_handle h);

extern "C" int dlra_create(int device, int64_t n_local, int64_t m, int r0, int rmax, int flags, dlra_handle* out) {
    if (!out) { g_create_error = "out == NULL"; return DLRA_EINVAL; }
    *out = nullptr;
    if (n_local < 1 || m < 1 || r0 < 1 || rmax < r0 || rmax > 128 || rmax > m) {
        g_create_error = "dlra_create: need n_local >= 1, m >= 1, 1 <= r0 <= rmax <= min(128, m)";
        return DLRA_EINVAL;
    }
    dlra_engine* e = new dlra_engine();
    try {
        int ndev = 0;
        DLRA_CUDA(cudaGetDeviceCount(&ndev));
        DLRA_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
        DLRA_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        DLRA_CUDA(cudaGetDeviceProperties(&prop, device));
        DLRA_REQUIRE(prop.major >= 10, "libdlra.so is built for sm_100a (B200) only");
        e->device = device;
        e->n = n_local; e->m = m; e->r = r0; e->rmax = rmax; e->flags = flags;
        e->W = (flags & DLRA_RANK_ADAPTIVE) ? 2 * rmax : rmax;
        e->cx.num_sms = prop.multiProcessorCount;
        // the main stream outranks the auxiliary one: when both have CTAs pending (the streaming pass and a small m-side / Gram
        // kernel that became ready at the same moment) the pass gets the SMs first
        int prio_least = 0, prio_greatest = 0;
        DLRA_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        DLRA_CUDA(cudaStreamCreateWithPriority(&e->cx.stream, cudaStreamNonBlocking, prio_greatest));
        DLRA_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        DLRA_CUDA(cudaStreamCreateWithPriority(&e->ax.stream, cudaStreamNonBlocking, prio_least));
        e->ax.num_sms = e->cx.num_sms;
        DLRA_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
        DLRA_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
        DLRA_CUDA(cudaEventCreateWithFlags(&e->ev_rnew, cudaEventDisableTiming));
        DLRA_CUDA(cudaEventCreateWithFlags(&e->ev_kqr, cudaEventDisableTiming));
        DLRA_CUDA(cudaEventCreateWithFlags(&e->ev_pass, cudaEventDisableTiming));
        DLRA_CUDA(cudaEventCreateWithFlags(&e->ev_lsum, cudaEventDisableTiming));
        DLRA_CUDA(cudaEventCreateWithFlags(&e->ev_join2, cudaEventDisableTiming));
        const int64_t W = e->W;
        auto dmalloc = [&](int64_t doubles) {
            double* p = nullptr;
            cudaError_t er = cudaMalloc(&p, (size_t)doubles * 8);
            if (er != cudaSuccess) throw CudaError(5, std::string("cudaMalloc failed: ") + cudaGetErrorString(er));
            DLRA_CUDA(cudaMemsetAsync(p, 0, (size_t)doubles * 8, e->cx.stream));
            return p;
        };
        e->U = dmalloc(e->n * W); e->UB = dmalloc(e->n * W);
        e->V = dmalloc(e->m * W); e->VB = dmalloc(e->m * W);
        e->small_block = dmalloc(18 * W * W + 64);
        double* sb = e->small_block;
        e->S = sb; sb += W * W; e->M = sb; sb += W * W; e->N = sb; sb += W * W; e->Sh = sb; sb += W * W;
        e->T1 = sb; sb += W * W; e->T2 = sb; sb += W * W; e->Rm = sb; sb += W * W; e->Pm = sb; sb += W * W;
        e->Qm = sb; sb += W * W; e->stg = sb; sb += 8 * W * W; e->sig = sb; sb += W; e->scal_dev = sb + W;
        DLRA_CUDA(cudaMalloc(&e->cx.counters, (512 + 16) * sizeof(unsigned int)));
        DLRA_CUDA(cudaMemsetAsync(e->cx.counters, 0, (512 + 16) * sizeof(unsigned int), e->cx.stream));
        e->ax.counters = e->cx.counters + 256;
        e->cx.sync = e->cx.counters + 512;
        e->ax.sync = e->cx.counters + 520;
        DLRA_CUDA(cudaMalloc(&e->r_new_dev, sizeof(int)));
        DLRA_CUDA(cudaHostAlloc(&e->r_new_host, sizeof(int), cudaHostAllocDefault));
        *e->r_new_host = r0;
        for (int i = 0; i < dlra_engine::NOWN; ++i) {
            DLRA_CUDA(cudaEventCreateWithFlags(&e->own_free[i], cudaEventDisableTiming));
            DLRA_CUDA(cudaEventCreateWithFlags(&e->own_ready[i], cudaEventDisableTiming));
        }
        e->sub[0] = SubStepperCfg(); e->sub[1] = SubStepperCfg(); e->sub[2] = SubStepperCfg();
        e->phase_timing = getenv("DLRA_PHASES") != nullptr;
        e->cx.tall_eng = e;
        if (!getenv("DLRA_NO_TALL_GEMM")) e->cx.tall_gemm = tall_gemm_hook;
        if (e->phase_timing) e->cx.mark_fn = phase_mark_hook;
        DLRA_CUDA(cudaStreamSynchronize(e->cx.stream));
    } catch (const CudaError& ex) {
        g_create_error = ex.what();
        int code = ex.code;
        dlra_destroy(e);   // releases whatever was allocated before the failure
        return code;
    } catch (const std::exception& ex) {
        g_create_error = ex.what();
        dlra_destroy(e);
        return DLRA_ECUDA;
    }
    *out = e;
    return DLRA_OK;
}

extern "C" int dlra_destroy(dlra_handle h) {
    if (!h) return DLRA_OK;
    cudaSetDevice(h->device);
    if (h->cx.stream) cudaStreamSynchronize(h->cx.stream);
    if (h->ax.stream) cudaStreamSynchronize(h->ax.stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    h->comm.destroy();
    cudaFree(h->U); cudaFree(h->UB); cudaFree(h->V); cudaFree(h->VB); cudaFree(h->small_block);
    cudaFree(h->r_new_dev); cudaFreeHost(h->r_new_host); cudaFree(h->cx.counters);
    for (int i = 0; i < dlra_engine::NOWN; ++i) {
        if (h->own[i]) cudaFree(h->own[i]);
        if (h->own_free[i]) cudaEventDestroy(h->own_free[i]);
        if (h->own_ready[i]) cudaEventDestroy(h->own_ready[i]);
    }
    h->gws.release(); h->tws.release(); h->wtmp.release(); h->jws.release(); h->nscr.release(); h->mscr.release(); h->part.release(); h->isvd.release(); h->bstage.release();
    for (auto& pr : h->pass_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto& pm : h->phase_marks) cudaEventDestroy(pm.second);
    for (int i = 0; i < dlra_engine::NPROG; ++i) if (h->prog_ev[i]) cudaEventDestroy(h->prog_ev[i]);
    for (int i = 0; i < 8; ++i) if (h->user_events[i]) cudaEventDestroy(h->user_events[i]);
    de_release(h);
    for (int i = 0; i < 2; ++i) {
        h->save_stage[i].release();
        if (h->save_staged[i]) cudaEventDestroy(h->save_staged[i]);
        if (h->save_landed[i]) cudaEventDestroy(h->save_landed[i]);
    }
    h->gws2.release(); h->tws2.release(); h->wtmp2.release(); h->zcarry.k.release();
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_rnew) cudaEventDestroy(h->ev_rnew);
    if (h->ev_kqr) cudaEventDestroy(h->ev_kqr);
    if (h->ev_pass) cudaEventDestroy(h->ev_pass);
    if (h->ev_lsum) cudaEventDestroy(h->ev_lsum);
    if (h->ev_join2) cudaEventDestroy(h->ev_join2);
    h->UC.release();
    if (h->ev_ext) cudaEventDestroy(h->ev_ext);
    if (h->cx.stream) cudaStreamDestroy(h->cx.stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ax.stream) cudaStreamDestroy(h->ax.stream);
    cudaGetLastError();   // a partially constructed handle may have produced sticky-free errors above
    delete h;
    return DLRA_OK;
}

extern "C" int dlra_sync(dlra_handle h) {
    DLRA_API_BEGIN(h)
    DLRA_CUDA(cudaStreamSynchronize(h->cx.stream));
    // the one-launch TSQR bounds its inter-CTA waits instead of hanging the device; a time-out (the CTAs of one launch were not
    // co-resident for seconds: device shared with a foreign long-running kernel) must not pass silently
    if (h->cx.sync) {
        unsigned int w[12] = {0};   // cx.sync[0..3 of 8], ax.sync = cx.sync + 8
        DLRA_CUDA(cudaMemcpy(w, h->cx.sync, sizeof(w), cudaMemcpyDeviceToHost));
        if (w[2] != 0 || w[10] != 0)
            throw CudaError(DLRA_ECUDA, "an inter-CTA wait of the one-launch TSQR timed out: factors computed since then are invalid");
    }
    DLRA_API_END(h)
}

extern "C" int dlra_wait_stream(dlra_handle h, void* producer_stream) {
    DLRA_API_BEGIN(h)
    if (!h->ev_ext) DLRA_CUDA(cudaEventCreateWithFlags(&h->ev_ext, cudaEventDisableTiming));
    DLRA_CUDA(cudaEventRecord(h->ev_ext, (cudaStream_t)producer_stream));
    DLRA_CUDA(cudaStreamWaitEvent(h->cx.stream, h->ev_ext, 0));
    DLRA_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_ext, 0));
    DLRA_API_END(h)
}

static void progress_poll(dlra_handle h, int64_t wait_for);
extern "C" int dlra_progress(dlra_handle h, int64_t* steps_enqueued, int64_t* steps_completed, int64_t wait_for) {
    DLRA_API_BEGIN(h)
    progress_poll(h, std::min(wait_for, h->steps_enqueued));
    if (steps_enqueued) *steps_enqueued = h->steps_enqueued;
    if (steps_completed) *steps_completed = h->steps_completed;
    DLRA_API_END(h)
}

extern "C" int dlra_get_stream(dlra_handle h, void** stream) {
    if (!h || !stream) return DLRA_EINVAL;
    *stream = (void*)h->cx.stream;
    return DLRA_OK;
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU
// ---------------------------------------------------------------------------------------------------
extern "C" int dlra_nccl_unique_id(void* id128) {
    try {
        Comm c;
        c.load();
        Comm::UniqueId id;
        c.check(c.pGetUniqueId(&id), "ncclGetUniqueId");
        memcpy(id128, id.internal, 128);
    } catch (const CudaError& ex) { g_create_error = ex.what(); return ex.code; }
    return DLRA_OK;
}

extern "C" int dlra_comm_init(dlra_handle h, int nranks, int rank, const void* id128) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks && id128, "bad communicator arguments");
    DLRA_REQUIRE(nranks <= 8, "row sharding supports up to 8 ranks (one NVSwitch box)");
    if (nranks > 1) h->comm.init(nranks, rank, id128);
    DLRA_API_END(h)
}

extern "C" int dlra_p2p_export(dlra_handle h, void* handle64) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(handle64 != nullptr, "null output");
    // largest message: an m x W block of L (plus slack for the small r x r / R-factor messages)
    h->comm.p2p_alloc(((size_t)h->m * h->W + 8 * (size_t)h->W * h->W + 1024) * sizeof(double));
    h->comm.p2p_export(handle64);
    DLRA_API_END(h)
}

extern "C" int dlra_p2p_import(dlra_handle h, int nranks, int rank, const void* handles) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(handles != nullptr, "null handles");
    DLRA_REQUIRE(!h->rhs.set || nranks == 1, "DE right-hand sides are single-GPU");
    h->comm.p2p_import(nranks, rank, handles);
    DLRA_API_END(h)
}

// ---------------------------------------------------------------------------------------------------
// factors
// ---------------------------------------------------------------------------------------------------
static void set_factors_impl(dlra_handle h, const double* U, int64_t ldu, const double* S, int64_t lds, const double* V, int64_t ldv,
                             int r, cudaMemcpyKind kind) {
    DLRA_REQUIRE(r >= 1 && r <= h->rmax, "rank outside [1, rmax]");
    DLRA_REQUIRE(U && S && V && ldu >= h->n && lds >= r && ldv >= h->m, "bad factor pointers / leading dimensions");
    cudaStream_t s = h->cx.stream;
    DLRA_CUDA(cudaMemcpy2DAsync(h->U, h->n * 8, U, ldu * 8, h->n * 8, r, kind, s));
    DLRA_CUDA(cudaMemcpy2DAsync(h->V, h->m * 8, V, ldv * 8, h->m * 8, r, kind, s));
    DLRA_CUDA(cudaMemcpy2DAsync(h->S, (size_t)h->W * 8, S, lds * 8, (size_t)r * 8, r, kind, s));
    h->r = r;
    h->kl_ready = false;   // a precomputed K/L pass belongs to the factors it was formed with
    h->zcarry.valid = false;
    h->basis_trusted = false;   // user-supplied bases: the next rank-adaptive step factors the whole augmented basis
    if (kind == cudaMemcpyHostToDevice) DLRA_CUDA(cudaStreamSynchronize(s));
}
extern "C" int dlra_set_factors_host(dlra_handle h, const double* U, int64_t ldu, const double* S, int64_t lds, const double* V,
                                     int64_t ldv, int r) {
    DLRA_API_BEGIN(h)
    set_factors_impl(h, U, ldu, S, lds, V, ldv, r, cudaMemcpyHostToDevice);
    DLRA_API_END(h)
}
extern "C" int dlra_set_factors(dlra_handle h, const double* U, int64_t ldu, const double* S, int64_t lds, const double* V, int64_t ldv,
                                int r) {
    DLRA_API_BEGIN(h)
    set_factors_impl(h, U, ldu, S, lds, V, ldv, r, cudaMemcpyDeviceToDevice);
    DLRA_API_END(h)
}
static void get_factors_impl(dlra_handle h, double* U, int64_t ldu, double* S, int64_t lds, double* V, int64_t ldv, int* r,
                             cudaMemcpyKind kind) {
    const int rr = h->r;
    DLRA_REQUIRE(U && S && V && ldu >= h->n && lds >= rr && ldv >= h->m, "bad factor pointers / leading dimensions");
    cudaStream_t s = h->cx.stream;
    DLRA_CUDA(cudaMemcpy2DAsync(U, ldu * 8, h->U, h->n * 8, h->n * 8, rr, kind, s));
    DLRA_CUDA(cudaMemcpy2DAsync(V, ldv * 8, h->V, h->m * 8, h->m * 8, rr, kind, s));
    DLRA_CUDA(cudaMemcpy2DAsync(S, lds * 8, h->S, (size_t)h->W * 8, (size_t)rr * 8, rr, kind, s));
    DLRA_CUDA(cudaStreamSynchronize(s));
    if (r) *r = rr;
}
extern "C" int dlra_get_factors_host(dlra_handle h, double* U, int64_t ldu, double* S, int64_t lds, double* V, int64_t ldv, int* r) {
    DLRA_API_BEGIN(h)
    get_factors_impl(h, U, ldu, S, lds, V, ldv, r, cudaMemcpyDeviceToHost);
    DLRA_API_END(h)
}
extern "C" int dlra_get_factors(dlra_handle h, double* U, int64_t ldu, double* S, int64_t lds, double* V, int64_t ldv, int* r) {
    DLRA_API_BEGIN(h)
    get_factors_impl(h, U, ldu, S, lds, V, ldv, r, cudaMemcpyDeviceToDevice);
    DLRA_API_END(h)
}
// update_sol! without stalling the step stream (primitives.jl:82-90 deep-copies u after every step): the factors are copied
// device-to-device into one of two staging slots ON the compute stream (a few µs), and the copy stream moves the slot to the
// caller's (pinned) host buffers while the next steps run.  The compute stream only waits if both slots are still in flight.
extern "C" int dlra_save_factors_async(dlra_handle h, double* U, int64_t ldu, double* S, int64_t lds, double* V, int64_t ldv, int* r) {
    DLRA_API_BEGIN(h)
    const int rr = h->r;
    DLRA_REQUIRE(U && S && V && ldu >= h->n && lds >= rr && ldv >= h->m, "bad factor pointers / leading dimensions");
    const int slot = h->save_next;
    h->save_next ^= 1;
    if (!h->save_staged[slot]) {
        DLRA_CUDA(cudaEventCreateWithFlags(&h->save_staged[slot], cudaEventDisableTiming));
        DLRA_CUDA(cudaEventCreateWithFlags(&h->save_landed[slot], cudaEventDisableTiming));
    } else {
        DLRA_CUDA(cudaStreamWaitEvent(h->cx.stream, h->save_landed[slot], 0));   // the previous D2H out of this slot is done
    }
    const int64_t W = h->W;
    h->save_stage[slot].ensure((h->n + h->m + W) * W, h->cx.stream);
    double* su = h->save_stage[slot].p;
    double* sv = su + h->n * W;
    double* ss = sv + h->m * W;
    cudaStream_t cs = h->cx.stream, ds = h->copy_stream;
    DLRA_CUDA(cudaMemcpyAsync(su, h->U, (size_t)h->n * rr * 8, cudaMemcpyDeviceToDevice, cs));
    DLRA_CUDA(cudaMemcpyAsync(sv, h->V, (size_t)h->m * rr * 8, cudaMemcpyDeviceToDevice, cs));
    DLRA_CUDA(cudaMemcpy2DAsync(ss, (size_t)rr * 8, h->S, (size_t)W * 8, (size_t)rr * 8, rr, cudaMemcpyDeviceToDevice, cs));
    DLRA_CUDA(cudaEventRecord(h->save_staged[slot], cs));
    DLRA_CUDA(cudaStreamWaitEvent(ds, h->save_staged[slot], 0));
    DLRA_CUDA(cudaMemcpy2DAsync(U, ldu * 8, su, h->n * 8, h->n * 8, rr, cudaMemcpyDeviceToHost, ds));
    DLRA_CUDA(cudaMemcpy2DAsync(V, ldv * 8, sv, h->m * 8, h->m * 8, rr, cudaMemcpyDeviceToHost, ds));
    DLRA_CUDA(cudaMemcpy2DAsync(S, lds * 8, ss, (size_t)rr * 8, (size_t)rr * 8, rr, cudaMemcpyDeviceToHost, ds));
    DLRA_CUDA(cudaEventRecord(h->save_landed[slot], ds));
    if (r) *r = rr;
    DLRA_API_END(h)
}
extern "C" int dlra_save_wait(dlra_handle h) {
    DLRA_API_BEGIN(h)
    for (int i = 0; i < 2; ++i) if (h->save_landed[i]) DLRA_CUDA(cudaEventSynchronize(h->save_landed[i]));
    DLRA_API_END(h)
}
extern "C" int dlra_get_rank(dlra_handle h, int* r) {
    if (!h || !r) return DLRA_EINVAL;
    *r = h->r;
    return DLRA_OK;
}
extern "C" int dlra_factor_ptrs(dlra_handle h, const double** U, int64_t* ldu, const double** S, int64_t* lds, const double** V,
                                int64_t* ldv, int* r) {
    if (!h) return DLRA_EINVAL;
    if (U) *U = h->U; if (ldu) *ldu = h->n;
    if (S) *S = h->S; if (lds) *lds = h->W;
    if (V) *V = h->V; if (ldv) *ldv = h->m;
    if (r) *r = h->r;
    return DLRA_OK;
}

// ---------------------------------------------------------------------------------------------------
// data feed
// ---------------------------------------------------------------------------------------------------
static int own_slot_for_copy(dlra_handle h) {
    // rotate over four engine-owned n x m buffers (prev, cur, lookahead + one in flight) so that H2D copies overlap the steps
    int slot = -1;
    for (int tries = 0; tries < dlra_engine::NOWN; ++tries) {
        int cand = (h->own_next + tries) % dlra_engine::NOWN;
        if (cand != h->cur_own && cand != h->prev_own && cand != h->nxt_own) { slot = cand; break; }
    }
    DLRA_REQUIRE(slot >= 0, "no free snapshot buffer");
    h->own_next = (slot + 1) % dlra_engine::NOWN;
    if (!h->own[slot]) {
        cudaError_t er = cudaMalloc(&h->own[slot], (size_t)h->n * h->m * 8);
        if (er != cudaSuccess) throw CudaError(5, std::string("cudaMalloc(snapshot buffer) failed: ") + cudaGetErrorString(er));
        DLRA_CUDA(cudaEventRecord(h->own_free[slot], h->cx.stream));
    }
    return slot;
}
static void copy_host_snapshot(dlra_handle h, int slot, const double* A, int64_t ld) {
    DLRA_CUDA(cudaStreamWaitEvent(h->copy_stream, h->own_free[slot], 0));
    DLRA_CUDA(cudaMemcpy2DAsync(h->own[slot], h->n * 8, A, ld * 8, h->n * 8, h->m, cudaMemcpyHostToDevice, h->copy_stream));
    DLRA_CUDA(cudaEventRecord(h->own_ready[slot], h->copy_stream));
}

extern "C" int dlra_data_init(dlra_handle h, const double* A0, int64_t ld) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(A0 && ld >= h->n, "bad snapshot pointer / leading dimension");
    h->prev = A0; h->ldprev = ld; h->prev_own = -1; h->have_cur = false; h->cur_own = -1;
    h->have_nxt = false; h->nxt_own = -1; h->kl_ready = false;
    DLRA_API_END(h)
}
extern "C" int dlra_data_init_host(dlra_handle h, const double* A0, int64_t ld) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(A0 && ld >= h->n, "bad snapshot pointer / leading dimension");
    h->prev_own = -1; h->cur_own = -1; h->have_cur = false; h->have_nxt = false; h->nxt_own = -1; h->kl_ready = false;
    int slot = own_slot_for_copy(h);
    copy_host_snapshot(h, slot, A0, ld);
    DLRA_CUDA(cudaStreamWaitEvent(h->cx.stream, h->own_ready[slot], 0));
    h->prev = h->own[slot]; h->ldprev = h->n; h->prev_own = slot;
    DLRA_API_END(h)
}
extern "C" int dlra_data_push(dlra_handle h, const double* A, int64_t ld, int kind) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(A && ld >= h->n, "bad snapshot pointer / leading dimension");
    DLRA_REQUIRE(kind == DLRA_DATA_SNAPSHOT || kind == DLRA_DATA_DELTA, "bad data kind");
    if (h->have_cur) {   // second push before the step: one-snapshot lookahead (enables the software-pipelined BUG pass)
        DLRA_REQUIRE(!h->have_nxt, "at most one snapshot of lookahead can be pushed");
        h->nxt = A; h->ldnxt = ld; h->nxt_kind = kind; h->have_nxt = true; h->nxt_own = -1;
    } else {
        h->cur = A; h->ldcur = ld; h->cur_kind = kind; h->have_cur = true; h->cur_own = -1;
    }
    DLRA_API_END(h)
}
extern "C" int dlra_data_push_host(dlra_handle h, const double* A, int64_t ld, int kind) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(A && ld >= h->n, "bad snapshot pointer / leading dimension");
    DLRA_REQUIRE(kind == DLRA_DATA_SNAPSHOT || kind == DLRA_DATA_DELTA, "bad data kind");
    DLRA_REQUIRE(!(h->have_cur && h->have_nxt), "at most one snapshot of lookahead can be pushed");
    int slot = own_slot_for_copy(h);
    copy_host_snapshot(h, slot, A, ld);
    if (h->have_cur) { h->nxt = h->own[slot]; h->ldnxt = h->n; h->nxt_kind = kind; h->have_nxt = true; h->nxt_own = slot; }
    else { h->cur = h->own[slot]; h->ldcur = h->n; h->cur_kind = kind; h->have_cur = true; h->cur_own = slot; }
    DLRA_API_END(h)
}

// Fetch the increment for this step (A.0 of SURVEY.md Appendix A) and hand back what to do afterwards.
static Delta begin_data_step(dlra_handle h, bool want_full_snapshot) {
    DLRA_REQUIRE(h->have_cur, "no data pushed for this step (call dlra_data_push[_host] first)");
    Delta d;
    d.A = h->cur; d.lda = h->ldcur;
    if (h->cur_own >= 0) DLRA_CUDA(cudaStreamWaitEvent(h->cx.stream, h->own_ready[h->cur_own], 0));
    if (want_full_snapshot) {
        DLRA_REQUIRE(h->cur_kind == DLRA_DATA_SNAPSHOT, "the greedy step needs the full snapshot (DLRA_DATA_SNAPSHOT)");
    } else if (h->cur_kind == DLRA_DATA_SNAPSHOT) {
        DLRA_REQUIRE(h->prev != nullptr, "dlra_data_init[_host] must provide the initial snapshot before SNAPSHOT pushes");
        d.Aprev = h->prev; d.ldap = h->ldprev;
    }
    return d;
}
static void end_data_step(dlra_handle h) {
    // yprev .= ycurr  (projector_splitting.jl:121): a pointer rotation here
    if (h->cur_kind == DLRA_DATA_SNAPSHOT) {
        if (h->prev_own >= 0) DLRA_CUDA(cudaEventRecord(h->own_free[h->prev_own], h->cx.stream));
        h->prev = h->cur; h->ldprev = h->ldcur; h->prev_own = h->cur_own;
    } else if (h->cur_own >= 0) {
        DLRA_CUDA(cudaEventRecord(h->own_free[h->cur_own], h->cx.stream));
    }
    h->have_cur = false; h->cur_own = -1; h->cur = nullptr;
    if (h->have_nxt) {   // the lookahead snapshot becomes the pushed data of the next step
        h->cur = h->nxt; h->ldcur = h->ldnxt; h->cur_kind = h->nxt_kind; h->cur_own = h->nxt_own; h->have_cur = true;
        h->have_nxt = false; h->nxt = nullptr; h->nxt_own = -1;
    }
}

// ---------------------------------------------------------------------------------------------------
// shared step pieces
// ---------------------------------------------------------------------------------------------------
// scratch set of a stream: main (n-side and everything sequential) or auxiliary (replicated m-side chain)
struct Side {
    Ctx* cx; DevBuf* tws; DevBuf* gws; DevBuf* wtmp;
};
static Side main_side(dlra_handle h) { return Side{&h->cx, &h->tws, &h->gws, &h->wtmp}; }
static Side aux_side(dlra_handle h) { return Side{&h->ax, &h->tws2, &h->gws2, &h->wtmp2}; }
// the auxiliary stream starts after everything enqueued so far on the main stream ...
static void fork_aux(dlra_handle h) {
    DLRA_CUDA(cudaEventRecord(h->ev_fork, h->cx.stream));
    DLRA_CUDA(cudaStreamWaitEvent(h->ax.stream, h->ev_fork, 0));
}
// ... and the main stream continues once the auxiliary chain is done
static void join_aux(dlra_handle h) {
    DLRA_CUDA(cudaEventRecord(h->ev_join, h->ax.stream));
    DLRA_CUDA(cudaStreamWaitEvent(h->cx.stream, h->ev_join, 0));
}
// The pipelined BUG step leaves the sum of the next step's L partials running on the auxiliary stream; it writes VB.  A following BUG
// step consumes it in stream order on that same stream; every other user of VB on the main stream must wait for it first.
static void settle_aux(dlra_handle h) {
    if (!h->lsum_pending) return;
    DLRA_CUDA(cudaStreamWaitEvent(h->cx.stream, h->ev_lsum, 0));
    h->lsum_pending = false;
}
static void ensure_qr_ws(Side sd, int64_t rows, int C) {
    const int cb = std::min(C, TSQR_MAXC);
    sd.tws->ensure(tsqr_ws_size(rows, cb, 8), sd.cx->stream);
    if (C > TSQR_MAXC) {
        sd.wtmp->ensure(thin_qr_wtmp(C), sd.cx->stream);
        sd.gws->ensure(gemm_tn_ws(*sd.cx, rows, C, TSQR_MAXC), sd.cx->stream);
    }
}
// thin QR of an n-side (row sharded) matrix, in place
static void qr_nside(dlra_handle h, double* A, int C, double* R, int ortho_cols = 0) {
    NvtxRange nvtx_qr("dlra:qr_nside");
    Side sd = main_side(h);
    ensure_qr_ws(sd, h->n, C);
    thin_qr(h->cx, h->comm, h->n, C, A, h->n, A, h->n, R, h->W, h->tws.p, h->gws.p, h->wtmp.p, ortho_cols);
}
// thin QR of (A + Ua*Sa) in place; the rank-r update rides on the panel load of the first TSQR level when it can
static void qr_nside_plus(dlra_handle h, double* A, int C, const double* Ua, const double* Sa, int k) {
    NvtxRange nvtx_qr("dlra:qr_nside_plus");
    if (C <= TSQR_MAXC && h->n > 128) {
        Side sd = main_side(h);
        ensure_qr_ws(sd, h->n, C);
        TsqrAdd add; add.U = Ua; add.ldu = h->n; add.S = Sa; add.lds = h->W; add.k = k;
        tsqr(h->cx, h->comm, h->n, C, A, h->n, A, h->n, nullptr, h->W, h->tws.p, add);
    } else {
        gemm_nn(h->cx, h->n, k, C, Ua, h->n, nullptr, 0, Sa, h->W, false, A, h->n, 1.0, 1.0);
        qr_nside(h, A, C, nullptr);
    }
}
// thin QR of an m-side (replicated) matrix, in place, computed redundantly on every rank
static void qr_mside(dlra_handle h, Side sd, double* A, int C, double* R, int ortho_cols = 0) {
    NvtxRange nvtx_qr("dlra:qr_mside");
    ensure_qr_ws(sd, h->m, C);
    thin_qr(*sd.cx, h->self, h->m, C, A, h->m, A, h->m, R, h->W, sd.tws->p, sd.gws->p, sd.wtmp->p, ortho_cols);
}
// thin QR of (A + Va*Sa') in place on the m-side; the rank-k update rides on the panel load of the first TSQR level when it can
static void qr_mside_plus(dlra_handle h, Side sd, double* A, int C, const double* Va, const double* Sa, int k) {
    NvtxRange nvtx_qr("dlra:qr_mside_plus");
    if (C <= TSQR_MAXC && h->m > 128) {
        ensure_qr_ws(sd, h->m, C);
        TsqrAdd add; add.U = Va; add.ldu = h->m; add.S = Sa; add.lds = h->W; add.k = k; add.transS = true;
        tsqr(*sd.cx, h->self, h->m, C, A, h->m, A, h->m, nullptr, h->W, sd.tws->p, add);
    } else {
        gemm_nn(*sd.cx, h->m, k, C, Va, h->m, nullptr, 0, Sa, h->W, true, A, h->m, 1.0, 1.0);
        qr_mside(h, sd, A, C, nullptr);
    }
}
// p x q small matrix with ld W, summed over the row shards in place
static void allreduce_small(dlra_handle h, double* C, int p, int q) {
    if (h->comm.nranks <= 1) return;
    SmallMats sm;
    sm.n = 1;
    sm.p[0] = C; sm.rows[0] = p; sm.cols[0] = q; sm.ld[0] = h->W;
    h->comm.allreduce_small_mats(sm, h->stg, h->cx);
}
// C (p x q, ld W) = A' * B over the sharded n dimension (+ all-reduce)
static void gram_nside(dlra_handle h, int p, int q, const double* A, const double* B, double* C) {
    h->gws.ensure(gemm_tn_ws(h->cx, h->n, p, q), h->cx.stream);
    gemm_tn(h->cx, h->n, p, q, A, h->n, nullptr, 0, B, h->n, C, h->W, 1.0, 0.0, h->gws.p);
    allreduce_small(h, C, p, q);
}
// local (this rank's rows) part of A'*B: the cross-rank sum is deferred and merged with a later collective
static void gram_nside_local(dlra_handle h, int p, int q, const double* A, const double* B, double* C) {
    h->gws.ensure(gemm_tn_ws(h->cx, h->n, p, q), h->cx.stream);
    gemm_tn(h->cx, h->n, p, q, A, h->n, nullptr, 0, B, h->n, C, h->W, 1.0, 0.0, h->gws.p);
}
// one all-reduce for two small matrices (ld W): halves the number of cross-GPU synchronisation points per step
static void allreduce_pair(dlra_handle h, double* A, int pa, int qa, double* B, int pb, int qb) {
    if (h->comm.nranks <= 1) return;
    SmallMats sm;
    sm.n = 2;
    sm.p[0] = A; sm.rows[0] = pa; sm.cols[0] = qa; sm.ld[0] = h->W;
    sm.p[1] = B; sm.rows[1] = pb; sm.cols[1] = qb; sm.ld[1] = h->W;
    h->comm.allreduce_small_mats(sm, h->stg /* 8*W*W doubles */, h->cx);   // P2P transport: one single-CTA launch
}
static void gram_mside(dlra_handle h, Side sd, int p, int q, const double* A, const double* B, double* C) {
    sd.gws->ensure(gemm_tn_ws(*sd.cx, h->m, p, q), sd.cx->stream);
    gemm_tn(*sd.cx, h->m, p, q, A, h->m, nullptr, 0, B, h->m, C, h->W, 1.0, 0.0, sd.gws->p);
}

// K-flow / L-flow / S-flow of one outer step: data problems contract the increment, DE problems integrate
// the projected right-hand side (de_flows.cuh).  `Kbuf` holds K0 on entry and K1 on exit, etc.
struct StepCtx {
    bool is_data;
    Delta d;
    double t, dt;
};

// ---------------------------------------------------------------------------------------------------
// unconventional (BUG) step — unconventional.jl:133-157, SURVEY.md A.2
// ---------------------------------------------------------------------------------------------------
static void bug_step(dlra_handle h, const StepCtx& sc) {
    NvtxRange nvtx_step("dlra:bug_step");
    Ctx& cx = h->cx;
    const int r = h->r;
    const int64_t n = h->n, m = h->m, W = h->W;
    double *K = h->UB, *L = h->VB;
    // K = U0*S0 (+ ΔA*V0);  L = V0*S0' (+ ΔA'*U0)   — one fused read of ΔA
    const bool pre = sc.is_data && h->kl_ready && h->kl_rank == r;   // formed by the previous step's pipelined pass
    const bool lsum_done = pre && h->kl_lsum_done;                  // ... whose L partials were already summed into VB
    h->kl_ready = false;
    h->kl_lsum_done = false;
    if (lsum_done) h->lsum_pending = false;   // consumed on the auxiliary stream, in stream order behind the sum
    else settle_aux(h);
    phase_mark(h, "step");
    // Single GPU: the whole L-side chain runs beside the K-side chain.  Row-sharded runs keep the cross-rank sum of L on the main
    // stream: measured at N = 2 (profiles/r02/multi_gpu_phases.txt) the spinning exchange kernel on the auxiliary stream competes
    // with the K-side TSQR for SMs and sets up a ~120 us ping-pong of arrival skew between the ranks (DLRA_LFIN_AUX=1 re-enables it).
    static const bool lfin_aux_multi = getenv("DLRA_LFIN_AUX") != nullptr;
    const bool lfin_aux = pre && !lsum_done && (h->comm.nranks <= 1 || (h->comm.p2p && lfin_aux_multi));
    if (pre) {
        // K = ΔA*V0 (already in UB) + U0*S0: the update is folded into the TSQR panel load below
        if (!lfin_aux && !lsum_done)
            l_finalize(h, r, h->kl_nparts, h->part.p, h->kl_ldlp, 16, 1, h->V, m, h->S, W, r, L, m);  // L = ΔA'*U0 + V0*S0'
    } else {
        gemm_nn(cx, n, r, r, h->U, n, nullptr, 0, h->S, W, false, K, n, 1.0, 0.0);
    }
    if (pre) {
    } else if (sc.is_data) {
        pass_KL(h, sc.d, r, h->V, m, h->U, n, K, n, L, m, h->V, m, h->S, W);   // L complete: all-reduced, + V0*S0'
    } else {
        gemm_nn(cx, m, r, r, h->V, m, nullptr, 0, h->S, W, true, L, m, 1.0, 0.0);
        de_K_flow(h, K, r, h->V, sc.t, sc.dt);
        de_L_flow(h, L, r, h->U, sc.t, sc.dt);
    }
    phase_mark(h, "KL_or_Lsum");
    const bool pipe = sc.is_data && !(h->flags & DLRA_FORCE_GENERIC) && h->have_nxt && h->nxt_kind == DLRA_DATA_SNAPSHOT &&
                      sc.d.Aprev != nullptr && r <= 16 && tma_pass_supported(n, m, sc.d) && tma_ok(h->nxt, h->ldnxt);
    // Pipelined pass: M = U1'*U0 (local rows) leaves the critical path.  It runs on the auxiliary stream BESIDE the streaming pass
    // (which then writes the next step's K into a third buffer instead of the buffer of U0) and is only needed by the core update /
    // the M + core all-reduce after the pass.  No exchange runs on the auxiliary stream because of it.
    static const bool gram_m_aux = !(getenv("DLRA_GRAM_M_AUX") && atoi(getenv("DLRA_GRAM_M_AUX")) == 0);
    static const bool gram_m_aux_multi = !(getenv("DLRA_GRAM_M_AUX_MULTI") && atoi(getenv("DLRA_GRAM_M_AUX_MULTI")) == 0);
    const bool m_aux = pipe && gram_m_aux && (h->comm.nranks <= 1 || gram_m_aux_multi);
    fork_aux(h);                                          // m-side chain on the auxiliary stream ...
    if (lfin_aux) l_finalize(h, r, h->kl_nparts, h->part.p, h->kl_ldlp, 16, 1, h->V, m, h->S, W, r, L, m, &h->ax);
    if (lsum_done) qr_mside_plus(h, aux_side(h), L, r, h->V, h->S, r);   // V1 = qr(ΔA'*U0 + V0*S0').Q, the update folded into the panel load
    else qr_mside(h, aux_side(h), L, r, nullptr);         // V1 = qr(L).Q
    gram_mside(h, aux_side(h), r, r, L, h->V, h->N);      // N = V1'*V0
    if (m_aux) DLRA_CUDA(cudaEventRecord(h->ev_join, h->ax.stream));   // the m-side is complete here
    if (pre) qr_nside_plus(h, K, r, h->U, h->S, r);       // ... overlaps U1 = qr(ΔA*V0 + U0*S0).Q
    else qr_nside(h, K, r, nullptr);                      //              U1 = qr(K).Q
    phase_mark(h, "qr_K");
    if (m_aux) {
        h->UC.ensure(n * W, cx.stream);
        DLRA_CUDA(cudaEventRecord(h->ev_kqr, cx.stream));
        DLRA_CUDA(cudaStreamWaitEvent(h->ax.stream, h->ev_kqr, 0));
        h->gws2.ensure(gemm_tn_ws(h->ax, n, r, r), h->ax.stream);
        gemm_tn(h->ax, n, r, r, K, n, nullptr, 0, h->U, n, h->M, W, 1.0, 0.0, h->gws2.p);   // M = U1'*U0 beside the pass
        DLRA_CUDA(cudaEventRecord(h->ev_join2, h->ax.stream));
        DLRA_CUDA(cudaStreamWaitEvent(cx.stream, h->ev_join, 0));
    } else {
        // M must be formed BEFORE the pipelined pass: that pass writes the next step's K into the buffer of U0
        gram_nside_local(h, r, r, K, h->U, h->M);         // M = U1'*U0 (local rows; summed over ranks below)
        phase_mark(h, "gram_M");
        join_aux(h);
    }
    phase_mark(h, "join_Lside");
    if (sc.is_data) {
        if (pipe) {
            // one sweep over A(t), A(t+dt), A(t+2dt): W = ΔA_k*V1 (this step's core) and ΔA_{k+1}*V1, ΔA_{k+1}'*U1 (next step)
            if (h->nxt_own >= 0) DLRA_CUDA(cudaStreamWaitEvent(cx.stream, h->own_ready[h->nxt_own], 0));
            const int nsub = choose_nsub(n, cx.num_sms);
            const int npanels = (int)cdiv(cdiv(n, PT_SI), nsub);
            const int nparts = std::min(npanels, cx.num_sms);
            const int64_t ldlp = round_up(m, 2);
            h->part.ensure((int64_t)nparts * ldlp * 16, cx.stream);
            h->nscr.ensure(n * (int64_t)r, cx.stream);
            double* Knext = m_aux ? h->UC.p : h->U;   // the next step's K: third buffer, or the (no longer needed) buffer of U0
            tri_pass_launch(h, h->nxt, h->ldnxt, sc.d.A, sc.d.lda, sc.d.Aprev, sc.d.ldap, r, L, m, K, n, h->nscr.p, n,
                            Knext, n, h->part.p, ldlp, nsub, npanels);
            static const bool lsum_early = !(getenv("DLRA_LSUM_EARLY") && atoi(getenv("DLRA_LSUM_EARLY")) == 0);
            if (m_aux && lsum_early && m > 128 && h->comm.nranks <= 1) {
                // the next step's L = sum of the per-CTA partials (+ V1*S1', folded into its QR): the sum does not need S1, so it runs on
                // the auxiliary stream beside the core update instead of in front of the next step's m-side QR.  Target: the buffer of
                // V0 (dead once N = V1'V0 exists; it is the next step's L buffer after the swap below).
                DLRA_CUDA(cudaEventRecord(h->ev_pass, cx.stream));
                DLRA_CUDA(cudaStreamWaitEvent(h->ax.stream, h->ev_pass, 0));
                l_finalize(h, r, nparts, h->part.p, ldlp, 16, 1, nullptr, 0, nullptr, 0, r, h->V, m, &h->ax);
                DLRA_CUDA(cudaEventRecord(h->ev_lsum, h->ax.stream));
                h->kl_lsum_done = true;
                h->lsum_pending = true;
            }
            if (m_aux) DLRA_CUDA(cudaStreamWaitEvent(cx.stream, h->ev_join2, 0));   // M is needed from here on
            h->gws.ensure(std::max(gemm_tn_ws(cx, n, r, r), gram_core_ws(n)), cx.stream);
            static const bool fused_core = !(getenv("DLRA_FUSED_CORE") && atoi(getenv("DLRA_FUSED_CORE")) == 0);
            if (fused_core && h->comm.nranks <= 1 && r <= 16 && cx.counters && gram_core_ok(n)) {
                // Rm = U1'*W and S1 = M*S0*N' + Rm in one launch (same arithmetic as the two kernels below)
                h->kl_ready = true; h->kl_nparts = nparts; h->kl_ldlp = ldlp; h->kl_rank = r;
                gram_core(cx, n, r, K, n, h->nscr.p, n, h->Rm, h->gws.p, h->M, h->S, h->N, h->S, (int)W);
                phase_mark(h, "pass_S(+KL_next)+gram+core");
                std::swap(h->U, h->UB);                       // U = U1, UB = buffer of U0
                if (m_aux) std::swap(h->UB, h->UC.p);         // UB = next step's K, UC = buffer of U0 (free)
                std::swap(h->V, h->VB);
                return;
            }
            gemm_tn(cx, n, r, r, K, n, nullptr, 0, h->nscr.p, n, h->Rm, W, 1.0, 0.0, h->gws.p);   // Rm = U1'*W (local rows)
            h->kl_ready = true; h->kl_nparts = nparts; h->kl_ldlp = ldlp; h->kl_rank = r;
            phase_mark(h, "pass_S(+KL_next)+gram");
            if (h->comm.nranks > 1 && r <= 16 && h->comm.ll_fits(512)) {
                // all-reduce of M and the core increment + core update in one single-CTA launch
                ll_allreduce_core_kernel<<<1, 256, 0, cx.stream>>>(h->comm.next_ll(0), r, h->M, h->Rm, (int)W, h->S, h->N);
                cx.launches++;
                DLRA_CUDA(cudaGetLastError());
                phase_mark(h, "allreduce_M_S");
            } else {
                allreduce_pair(h, h->M, r, r, h->Rm, r, r);
                phase_mark(h, "allreduce_M_S");
                core_update(cx, r, r, r, h->M, h->S, h->N, h->Rm, h->S, (int)W, h->T1);
            }
            phase_mark(h, "core_update");
            std::swap(h->U, h->UB);
            if (m_aux) std::swap(h->UB, h->UC.p);
            std::swap(h->V, h->VB);
            return;
        } else {
            pass_S(h, sc.d, r, r, K, n, L, m, h->Rm, W);        // Rm = U1'*ΔA*V1 (local rows)
        }
        phase_mark(h, "pass_S(+KL_next)+gram");
        allreduce_pair(h, h->M, r, r, h->Rm, r, r);             // M and the core increment share one collective
        phase_mark(h, "allreduce_M_S");
        core_update(cx, r, r, r, h->M, h->S, h->N, h->Rm, h->S, (int)W, h->T1);   // S1 = M*S0*N' + U1'*ΔA*V1 (one launch)
        phase_mark(h, "core_update");
        std::swap(h->U, h->UB);
        std::swap(h->V, h->VB);
        return;
    } else {
        small_gemm(cx, r, r, r, h->M, (int)W, false, h->S, (int)W, false, h->T1, (int)W, 1.0, 0.0);
        small_gemm(cx, r, r, r, h->T1, (int)W, false, h->N, (int)W, true, h->Sh, (int)W, 1.0, 0.0);
        de_S_flow(h, h->Sh, r, r, K, L, +1.0, sc.t, sc.dt);
    }
    std::swap(h->U, h->UB);
    std::swap(h->V, h->VB);
    copy_mat(cx, r, r, h->Sh, W, false, h->S, W);
}

// ---------------------------------------------------------------------------------------------------
// projector splitting (KSL) — projector_splitting.jl:129-152 (primal), 166-189 (dual), SURVEY.md A.1
// ---------------------------------------------------------------------------------------------------
static void ksl_primal_step(dlra_handle h, const StepCtx& sc) {
    NvtxRange nvtx_step("dlra:ksl_primal_step");
    settle_aux(h);
    Ctx& cx = h->cx;
    h->kl_ready = false;
    const int r = h->r;
    const int64_t n = h->n, m = h->m, W = h->W;
    double *K = h->UB, *L = h->VB;
    gemm_nn(cx, n, r, r, h->U, n, nullptr, 0, h->S, W, false, K, n, 1.0, 0.0);               // K = U0*S0
    if (sc.is_data) pass_KL(h, sc.d, r, h->V, m, nullptr, 0, K, n, nullptr, 0);               //   += ΔA*V0
    else de_K_flow(h, K, r, h->V, sc.t, sc.dt);
    qr_nside(h, K, r, h->Rm);                                                                 // U1, R
    if (sc.is_data) {
        // W = ΔA'*U1 ; S~ = R − W'*V0 ; L = V0*S~' + W        (U1'ΔA V0 == W'V0)
        pass_KL(h, sc.d, r, nullptr, 0, K, n, nullptr, 0, L, m);
        gram_mside(h, main_side(h), r, r, L, h->V, h->T1);                                    // T1 = W'*V0
        copy_mat(cx, r, r, h->Rm, W, false, h->Sh, W);
        copy_mat(cx, r, r, h->T1, W, false, h->Sh, W, -1.0, 1.0);                             // Sh = R − W'V0
        gemm_nn(cx, m, r, r, h->V, m, nullptr, 0, h->Sh, W, true, L, m, 1.0, 1.0);            // L = W + V0*Sh'
    } else {
        copy_mat(cx, r, r, h->Rm, W, false, h->Sh, W);
        de_S_flow(h, h->Sh, r, r, K, h->V, -1.0, sc.t, sc.dt);                                // S-step with (U1, V0), minus sign
        gemm_nn(cx, m, r, r, h->V, m, nullptr, 0, h->Sh, W, true, L, m, 1.0, 0.0);            // L0 = V0*S~'
        de_L_flow(h, L, r, K, sc.t, sc.dt);
    }
    qr_mside(h, main_side(h), L, r, h->Rm);                                                   // V1, R_L
    copy_mat(cx, r, r, h->Rm, W, true, h->S, W);                                              // S1 = R_L'
    std::swap(h->U, h->UB);
    std::swap(h->V, h->VB);
}

static void ksl_dual_step(dlra_handle h, const StepCtx& sc) {
    NvtxRange nvtx_step("dlra:ksl_dual_step");
    settle_aux(h);
    Ctx& cx = h->cx;
    h->kl_ready = false;
    const int r = h->r;
    const int64_t n = h->n, m = h->m, W = h->W;
    double *K = h->UB, *L = h->VB;
    if (sc.is_data) {
        pass_KL(h, sc.d, r, nullptr, 0, h->U, n, nullptr, 0, L, m, h->V, m, h->S, W);         // L = ΔA'*U0 + V0*S0'
    } else {
        gemm_nn(cx, m, r, r, h->V, m, nullptr, 0, h->S, W, true, L, m, 1.0, 0.0);
        de_L_flow(h, L, r, h->U, sc.t, sc.dt);
    }
    qr_mside(h, main_side(h), L, r, h->Rm);                                                   // V1, R_L
    copy_mat(cx, r, r, h->Rm, W, true, h->Sh, W);                                             // Sh = R_L'
    if (sc.is_data) {
        // Wn = ΔA*V1 ; S~ = R_L' − U0'*Wn ; K = U0*S~ + Wn
        fill_mat(cx, n, r, K, n, 0.0, 0.0);
        pass_KL(h, sc.d, r, L, m, nullptr, 0, K, n, nullptr, 0);
        gram_nside(h, r, r, h->U, K, h->T1);                                                  // T1 = U0'*Wn
        copy_mat(cx, r, r, h->T1, W, false, h->Sh, W, -1.0, 1.0);
        gemm_nn(cx, n, r, r, h->U, n, nullptr, 0, h->Sh, W, false, K, n, 1.0, 1.0);
    } else {
        de_S_flow(h, h->Sh, r, r, h->U, L, -1.0, sc.t, sc.dt);                                // S-step with (U0, V1)
        gemm_nn(cx, n, r, r, h->U, n, nullptr, 0, h->Sh, W, false, K, n, 1.0, 0.0);           // K0 = U0*S~
        de_K_flow(h, K, r, L, sc.t, sc.dt);
    }
    qr_nside(h, K, r, h->Rm);                                                                 // U1, R
    copy_mat(cx, r, r, h->Rm, W, false, h->S, W);                                             // S1 = R
    std::swap(h->U, h->UB);
    std::swap(h->V, h->VB);
}

// ---------------------------------------------------------------------------------------------------
// rank-adaptive BUG — rank_adaptive_unconventional.jl:194-233 + alg_recache :133-169, SURVEY.md A.3
// ---------------------------------------------------------------------------------------------------
static void rabug_step(dlra_handle h, const StepCtx& sc, double tol, int64_t rcap64, int* r_new_out, int* changed) {
    NvtxRange nvtx_step("dlra:rabug_step");
    settle_aux(h);
    Ctx& cx = h->cx;
    const int r = h->r;
    const int r2 = 2 * r;
    const int64_t n = h->n, m = h->m, W = h->W;
    DLRA_REQUIRE(h->flags & DLRA_RANK_ADAPTIVE, "handle was created without DLRA_RANK_ADAPTIVE");
    h->kl_ready = false;
    DLRA_REQUIRE(r2 <= W, "augmented basis exceeds workspace");
    DLRA_REQUIRE(r2 <= m, "augmented basis wider than the matrix");
    int rcap = (int)std::min<int64_t>(rcap64, (int64_t)h->rmax);
    rcap = std::min(rcap, r2);
    phase_mark(h, "step");
    double *Kh = h->UB, *Lh = h->VB;   // [K U0] -> Uhat ; [L V0] -> Vhat
    // DLRA_AUG_BASIS_FIRST: [U0 K], [V0 L] — same span; the leading (orthonormal) panel then skips its factorisation
    const bool first = (h->flags & DLRA_AUG_BASIS_FIRST) != 0;
    const int ortho_lead = (first && h->basis_trusted && (h->aug_steps % dlra_engine::AUG_REORTHO_EVERY) != dlra_engine::AUG_REORTHO_EVERY - 1) ? r : 0;
    h->aug_steps++;
    double* Kc = first ? Kh + (int64_t)r * n : Kh;          // where K / L are formed
    double* Lc = first ? Lh + (int64_t)r * m : Lh;
    double* Uc = first ? Kh : Kh + (int64_t)r * n;          // where the copies of U0 / V0 go
    double* Vc = first ? Lh : Lh + (int64_t)r * m;
    gemm_nn(cx, n, r, r, h->U, n, nullptr, 0, h->S, W, false, Kc, n, 1.0, 0.0);
    if (sc.is_data) {
        pass_KL(h, sc.d, r, h->V, m, h->U, n, Kc, n, Lc, m, h->V, m, h->S, W);
    } else {
        gemm_nn(cx, m, r, r, h->V, m, nullptr, 0, h->S, W, true, Lc, m, 1.0, 0.0);
        de_K_flow(h, Kc, r, h->V, sc.t, sc.dt);
        de_L_flow(h, Lc, r, h->U, sc.t, sc.dt);
    }
    copy_mat(cx, n, r, h->U, n, false, Uc, n);             // Uhat[:, r+1:end] = U0 (reference order) | Uhat[:, 1:r] = U0
    copy_mat(cx, m, r, h->V, m, false, Vc, m);
    phase_mark(h, "KL_flows");
    fork_aux(h);
    qr_mside(h, aux_side(h), Lh, r2, nullptr, ortho_lead);
    gram_mside(h, aux_side(h), r2, r, Lh, h->V, h->N);                   // N = Vhat'*V0
    qr_nside(h, Kh, r2, nullptr, ortho_lead);
    gram_nside_local(h, r2, r, Kh, h->U, h->M);                          // M = Uhat'*U0 (2r x r), local rows
    phase_mark(h, "qr_Uhat+M");
    join_aux(h);
    phase_mark(h, "join_Vhat");
    if (sc.is_data) {
        pass_S(h, sc.d, r2, r2, Kh, n, Lh, m, h->Rm, W);
        allreduce_pair(h, h->M, r2, r, h->Rm, r2, r2);
        small_gemm(cx, r2, r, r, h->M, (int)W, false, h->S, (int)W, false, h->T1, (int)W, 1.0, 0.0);
        small_gemm(cx, r2, r2, r, h->T1, (int)W, false, h->N, (int)W, true, h->Sh, (int)W, 1.0, 0.0);
        copy_mat(cx, r2, r2, h->Rm, W, false, h->Sh, W, 1.0, 1.0);
    } else {
        small_gemm(cx, r2, r, r, h->M, (int)W, false, h->S, (int)W, false, h->T1, (int)W, 1.0, 0.0);
        small_gemm(cx, r2, r2, r, h->T1, (int)W, false, h->N, (int)W, true, h->Sh, (int)W, 1.0, 0.0);
        de_S_flow(h, h->Sh, r2, r2, Kh, Lh, +1.0, sc.t, sc.dt);
    }
    phase_mark(h, "S_flow_or_pass");
    h->jws.ensure((int64_t)jacobi_ws_doubles(r2), cx.stream);
    static const bool jacobi_pre = !(getenv("DLRA_JACOBI_PRE") && atoi(getenv("DLRA_JACOBI_PRE")) == 0);
    if (jacobi_pre && r2 >= 64) {   // measured: the small QR costs more than the sweeps it saves below 64 x 64
        // Shat = Q0*R (Householder, in place), then Jacobi on R' with the accumulator started at Q0 (jacobi.cuh)
        double* Rj = jacobi_ws_rfactor(h->jws.p, r2);
        Side sd = main_side(h);
        ensure_qr_ws(sd, r2, r2);
        thin_qr(cx, h->self, r2, r2, h->Sh, W, h->Sh, W, Rj, r2, sd.tws->p, sd.gws->p, sd.wtmp->p);
        phase_mark(h, "core_qr");
        jacobi_svd(cx, r2, Rj, r2, h->jws.p, h->Pm, (int)W, h->sig, h->Qm, (int)W, tol, rcap, h->r_new_dev, nullptr, h->Sh, (int)W);
    } else {
        jacobi_svd(cx, r2, h->Sh, (int)W, h->jws.p, h->Pm, (int)W, h->sig, h->Qm, (int)W, tol, rcap, h->r_new_dev, nullptr);
    }
    phase_mark(h, "core_svd");
    // The host needs the new rank (it sizes every later launch), but nothing that follows on the device does: the basis update runs
    // at the full candidate width rcap (columns beyond r1 are never read: the factor buffers are W wide) and is enqueued BEFORE the
    // host waits, so the device stays busy during the round trip and the host resumes as soon as the SVD — not the whole step — is done.
    DLRA_CUDA(cudaMemcpyAsync(h->r_new_host, h->r_new_dev, sizeof(int), cudaMemcpyDeviceToHost, cx.stream));
    DLRA_CUDA(cudaEventRecord(h->ev_rnew, cx.stream));
    gemm_nn(cx, n, r2, rcap, Kh, n, nullptr, 0, h->Pm, W, false, h->U, n, 1.0, 0.0);   // U1 = Uhat*P[:, 1:r1]
    gemm_nn(cx, m, r2, rcap, Lh, m, nullptr, 0, h->Qm, W, false, h->V, m, 1.0, 0.0);   // V1 = Vhat*Q[:, 1:r1]
    fill_mat(cx, rcap, rcap, h->S, W, 0.0, 0.0);
    copy_mat(cx, 1, rcap, h->sig, 1, false, h->S, W + 1);                              // S1 = Diagonal(sigma[1:r1]) (leading block)
    phase_mark(h, "new_factors");
    DLRA_CUDA(cudaEventSynchronize(h->ev_rnew));
    const int r1 = *h->r_new_host;
    DLRA_REQUIRE(r1 >= 1 && r1 <= rcap, "rank selection returned an impossible rank");
    h->basis_trusted = true;
    if (changed) *changed = (r1 != r) ? 1 : 0;
    if (r_new_out) *r_new_out = r1;
    h->r = r1;
    if (r1 != r) de_rank_changed(h);
}

// ---------------------------------------------------------------------------------------------------
// greedy re-projection — greedy_integrator.jl:94-104
// ---------------------------------------------------------------------------------------------------
static void greedy_step(dlra_handle h, const Delta& x) {
    NvtxRange nvtx_step("dlra:greedy_step");
    settle_aux(h);
    Ctx& cx = h->cx;
    h->kl_ready = false;
    const int r = h->r;
    const int64_t n = h->n, m = h->m, W = h->W;
    double *XV = h->UB, *XU = h->VB;
    fill_mat(cx, n, r, XV, n, 0.0, 0.0);
    pass_KL(h, x, r, h->V, m, h->U, n, XV, n, XU, m);        // XV = X*V ; XU = X'*U (all-reduced)
    fork_aux(h);
    qr_mside(h, aux_side(h), XU, r, nullptr);
    qr_nside(h, XV, r, nullptr);
    join_aux(h);
    pass_S(h, x, r, r, XV, n, XU, m, h->T1, W);              // S = U1'*X*V1
    allreduce_small(h, h->T1, r, r);
    copy_mat(cx, r, r, h->T1, W, false, h->S, W);
    std::swap(h->U, h->UB);
    std::swap(h->V, h->VB);
}

// Orthonormal completion of the left singular vectors of a small core (r x r, ld W), in place.  The one-sided Jacobi sweep never
// rotates column pairs below the rounding floor of the matrix, so the columns of P that belong to (numerically) zero singular values
// come out normalised but NOT orthogonal to the others, whereas LAPACK's svd (greedy_integrator.jl:88) always returns an orthonormal P.
// Householder QR of P with the signs of diag(R) folded back reproduces the converged leading columns to rounding (their R block is
// diag(+-1)) and replaces the rest by an orthonormal completion.  Uses Rm as scratch.
static void ortho_complete(dlra_handle h, double* P, int r) {
    Side sd = main_side(h);
    ensure_qr_ws(sd, r, r);
    thin_qr(h->cx, h->self, r, r, P, h->W, P, h->W, h->Rm, h->W, sd.tws->p, sd.gws->p, sd.wtmp->p);
    sign_fix_cols(h->cx, r, r, P, h->W, h->Rm, h->W);
}

// ---------------------------------------------------------------------------------------------------
// greedy step of the two-factor representation u = U*Z' — greedy_integrator.jl:72-92
// Z lives in the V slot, S stays at identity.
// ---------------------------------------------------------------------------------------------------
static void greedy_two_factor_step(dlra_handle h, const Delta& x, int mode, bool carry_fsal, double t, double dt) {
    NvtxRange nvtx_step("dlra:greedy_two_factor_step");
    settle_aux(h);
    Ctx& cx = h->cx;
    h->kl_ready = false;
    const int r = h->r;
    const int64_t n = h->n, m = h->m, W = h->W;
    double *XZ = h->UB, *Z = h->VB;
    if (mode == DLRA_GREEDY_DATA) {
        pass_KL(h, x, r, nullptr, 0, h->U, n, nullptr, 0, Z, m);                              // Z = X'*U (all-reduced)
    } else {
        DLRA_REQUIRE(h->comm.nranks == 1, "the hybrid Z-flow is single-GPU (projected operators are not reduced over row shards)");
        copy_mat(cx, m, r, h->V, m, false, Z, m);
        if (!carry_fsal) h->zcarry.valid = false;
        de_L_flow(h, Z, r, h->U, t, dt, &h->zcarry);                                          // dZ/dt = F(U Z')' U
    }
    fill_mat(cx, n, r, XZ, n, 0.0, 0.0);
    pass_KL(h, x, r, Z, m, nullptr, 0, XZ, n, nullptr, 0);                                    // XZ = X*Z
    // svd(XZ) = (Q_qr*P) * Sigma * Qj'  with  XZ = Q_qr*R,  R = P*Sigma*Qj'   =>   U = Q_qr * (P*Qj')
    qr_nside(h, XZ, r, h->Rm);
    h->jws.ensure((int64_t)jacobi_ws_doubles(r), cx.stream);
    jacobi_svd(cx, r, h->Rm, (int)W, h->jws.p, h->Pm, (int)W, h->sig, h->Qm, (int)W, 0.0, r, nullptr, nullptr);
    ortho_complete(h, h->Pm, r);   // rank(X*Z) < r: the polar factor must still be orthonormal
    small_gemm(cx, r, r, r, h->Pm, (int)W, false, h->Qm, (int)W, true, h->T1, (int)W, 1.0, 0.0);
    gemm_nn(cx, n, r, r, XZ, n, nullptr, 0, h->T1, W, false, h->U, n, 1.0, 0.0);              // mul!(U, Q, P')
    std::swap(h->V, h->VB);
    fill_mat(cx, r, r, h->S, W, 0.0, 1.0);
}

// ---------------------------------------------------------------------------------------------------
// step entry points
// ---------------------------------------------------------------------------------------------------
// A step that throws (sub-stepper maxiters, non-finite state, CUDA error) must leave the handle usable: U, S, V are only
// replaced at the very end of a step (pointer swap / final copy), so they still hold the factors before the step; this guard
// puts the adaptive controllers back, drops the precomputed K/L pass and re-joins the auxiliary stream.
static void progress_poll(dlra_handle h, int64_t wait_for) {
    while (h->steps_completed < h->steps_enqueued) {
        cudaEvent_t ev = h->prog_ev[(h->steps_completed + 1) % dlra_engine::NPROG];
        if (h->steps_completed < wait_for) DLRA_CUDA(cudaEventSynchronize(ev));
        else {
            cudaError_t q = cudaEventQuery(ev);
            if (q == cudaErrorNotReady) break;
            DLRA_CUDA(q);
        }
        h->steps_completed++;
    }
}
// every step entry point ends with this: event k marks the completion of step k on the compute stream
static void progress_mark(dlra_handle h) {
    if (h->steps_enqueued - h->steps_completed >= dlra_engine::NPROG - 1) progress_poll(h, h->steps_completed + 1);   // ring full
    const int64_t k = ++h->steps_enqueued;
    cudaEvent_t& ev = h->prog_ev[k % dlra_engine::NPROG];
    if (!ev) DLRA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    DLRA_CUDA(cudaEventRecord(ev, h->cx.stream));
}

struct StepGuard {
    dlra_handle h;
    SubStepperCfg saved[3];
    bool ok = false;
    explicit StepGuard(dlra_handle hh) : h(hh) { for (int i = 0; i < 3; ++i) saved[i] = h->sub[i]; }
    void commit() { ok = true; progress_mark(h); }
    ~StepGuard() {
        if (ok) return;
        for (int i = 0; i < 3; ++i) h->sub[i] = saved[i];
        h->kl_ready = false;
        h->zcarry.valid = false;
        cudaStreamSynchronize(h->ax.stream);
        cudaStreamSynchronize(h->cx.stream);
    }
};

static StepCtx make_ctx(dlra_handle h, double t, double dt) {
    StepCtx sc;
    sc.t = t; sc.dt = dt;
    sc.is_data = !h->rhs.set;
    if (sc.is_data) sc.d = begin_data_step(h, false);
    return sc;
}

extern "C" int dlra_step_ksl(dlra_handle h, int order, double t, double dt) {
    DLRA_API_BEGIN(h)
    StepGuard guard(h);
    DLRA_REQUIRE(order == DLRA_KSL_PRIMAL || order == DLRA_KSL_DUAL || order == DLRA_KSL_STRANG, "bad KSL order");
    if (order == DLRA_KSL_STRANG) {
        DLRA_REQUIRE(h->rhs.set, "Strang on a data problem needs two increments: push, PRIMAL(dt/2), push, DUAL(dt/2)");
        StepCtx a; a.is_data = false; a.t = t; a.dt = dt / 2;
        ksl_primal_step(h, a);
        StepCtx b; b.is_data = false; b.t = t + dt / 2; b.dt = dt / 2;
        ksl_dual_step(h, b);
    } else {
        StepCtx sc = make_ctx(h, t, dt);
        if (order == DLRA_KSL_PRIMAL) ksl_primal_step(h, sc); else ksl_dual_step(h, sc);
        if (sc.is_data) end_data_step(h);
    }
    guard.commit();
    DLRA_API_END(h)
}

extern "C" int dlra_step_bug(dlra_handle h, double t, double dt) {
    DLRA_API_BEGIN(h)
    StepGuard guard(h);
    StepCtx sc = make_ctx(h, t, dt);
    bug_step(h, sc);
    if (sc.is_data) end_data_step(h);
    guard.commit();
    DLRA_API_END(h)
}

extern "C" int dlra_step_rabug(dlra_handle h, double t, double dt, double tol, int64_t rmax, int* r_new, int* rank_changed) {
    DLRA_API_BEGIN(h)
    StepGuard guard(h);
    DLRA_REQUIRE(tol >= 0.0 && rmax >= 1, "bad tolerance / rank cap");
    StepCtx sc = make_ctx(h, t, dt);
    rabug_step(h, sc, tol, rmax, r_new, rank_changed);
    if (sc.is_data) end_data_step(h);
    guard.commit();
    DLRA_API_END(h)
}

extern "C" int dlra_step_greedy(dlra_handle h, double t, double dt) {
    (void)t; (void)dt;
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(!h->rhs.set, "the greedy integrator is defined for data problems");
    Delta x = begin_data_step(h, true);
    greedy_step(h, x);
    end_data_step(h);
    progress_mark(h);
    DLRA_API_END(h)
}

extern "C" int dlra_step_greedy_two_factor(dlra_handle h, int mode, int carry_fsal, double t, double dt) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(mode == DLRA_GREEDY_DATA || mode == DLRA_GREEDY_HYBRID, "bad greedy mode");
    if (mode == DLRA_GREEDY_DATA) DLRA_REQUIRE(!h->rhs.set, "DLRA_GREEDY_DATA is defined for data problems");
    else DLRA_REQUIRE(h->rhs.set, "DLRA_GREEDY_HYBRID needs the right-hand side of the Z-flow (dlra_rhs_set)");
    StepGuard guard(h);
    Delta x = begin_data_step(h, true);
    greedy_two_factor_step(h, x, mode, carry_fsal != 0, t, dt);
    end_data_step(h);
    guard.commit();
    DLRA_API_END(h)
}

// ---------------------------------------------------------------------------------------------------
// initial condition on the device
// ---------------------------------------------------------------------------------------------------
// uniform(-1,1) test matrix from a counter-based hash (splitmix64): identical on every rank for a given seed
__global__ void random_fill_kernel(int64_t rows, int cols, double* __restrict__ X, int64_t ldx, uint64_t seed) {
    const int64_t tot = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(e + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        X[(e % rows) + (e / rows) * ldx] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
}

extern "C" int dlra_truncated_svd(dlra_handle h, const double* A, int64_t ld, int r, double tol, int oversample, int power_iters,
                                  uint64_t seed) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(A && ld >= h->n, "bad matrix pointer / leading dimension");
    DLRA_REQUIRE(r >= 0 && r <= h->rmax && oversample >= 0 && power_iters >= 0, "bad rank / oversampling / iteration count");
    DLRA_REQUIRE(r > 0 || tol >= 0.0, "either a rank or a tolerance is needed");
    settle_aux(h);
    Ctx& cx = h->cx;
    const int64_t n = h->n, m = h->m;
    const int l = (int)std::min<int64_t>(std::min<int64_t>((r > 0 ? r : h->rmax) + oversample, 128), m);
    DLRA_REQUIRE(l >= std::max(r, 1), "sketch narrower than the requested rank");
    // temporaries: Y (n x l), Z (m x l), small l x l blocks
    h->isvd.ensure(n * l + m * l + 6 * (int64_t)l * l + l + 64, cx.stream);
    double* Y = h->isvd.p;
    double* Z = Y + n * l;
    double* Rb = Z + m * l;               // l x l, ld l
    double* Rt = Rb + (int64_t)l * l;     // Rb'
    double* P = Rt + (int64_t)l * l;
    double* Wm = P + (int64_t)l * l;
    double* sg = Wm + (int64_t)l * l;
    Delta d; d.A = A; d.lda = ld;
    Side sd = main_side(h);
    auto qr_cols = [&](double* X, int64_t rows, Comm& cm, double* R) {
        ensure_qr_ws(sd, rows, l);
        thin_qr(cx, cm, rows, l, X, rows, X, rows, R, l, h->tws.p, h->gws.p, h->wtmp.p);
    };
    random_fill_kernel<<<(unsigned)std::min<int64_t>(cdiv(m * l, 256), 1184), 256, 0, cx.stream>>>(m, l, Z, m, seed);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
    for (int it = 0; it <= power_iters; ++it) {
        DLRA_CUDA(cudaMemsetAsync(Y, 0, (size_t)n * l * sizeof(double), cx.stream));
        pass_KL(h, d, l, Z, m, nullptr, 0, Y, n, nullptr, 0);          // Y = A*Z
        qr_cols(Y, n, h->comm, nullptr);                                // Q = orth(Y)   (row sharded)
        pass_KL(h, d, l, nullptr, 0, Y, n, nullptr, 0, Z, m);          // Z = A'*Q      (all-reduced)
        if (it < power_iters) qr_cols(Z, m, h->self, nullptr);          // re-orthonormalise between power iterations
    }
    qr_cols(Z, m, h->self, Rb);                                         // A'Q = Qb*Rb  =>  A ~ Q*Rb'*Qb'
    copy_mat(cx, l, l, Rb, l, true, Rt, l);
    h->jws.ensure((int64_t)jacobi_ws_doubles(l), cx.stream);
    const int rcap = std::min(h->rmax, l);
    jacobi_svd(cx, l, Rt, l, h->jws.p, P, l, sg, Wm, l, tol, rcap, h->r_new_dev, nullptr);   // Rb' = P*diag(sg)*Wm'
    int rr = r;
    if (r == 0) {
        DLRA_CUDA(cudaMemcpyAsync(h->r_new_host, h->r_new_dev, sizeof(int), cudaMemcpyDeviceToHost, cx.stream));
        DLRA_CUDA(cudaStreamSynchronize(cx.stream));
        rr = *h->r_new_host;
    }
    DLRA_REQUIRE(rr >= 1 && rr <= h->rmax, "selected rank outside [1, rmax]");
    gemm_nn(cx, n, l, rr, Y, n, nullptr, 0, P, l, false, h->U, n, 1.0, 0.0);     // U = Q*P[:, 1:r]
    gemm_nn(cx, m, l, rr, Z, m, nullptr, 0, Wm, l, false, h->V, m, 1.0, 0.0);    // V = Qb*W[:, 1:r]
    fill_mat(cx, rr, rr, h->S, h->W, 0.0, 0.0);
    copy_mat(cx, 1, rr, sg, 1, false, h->S, h->W + 1);                           // S = Diagonal(sigma[1:r])
    h->r = rr;
    h->kl_ready = false;
    DLRA_API_END(h)
}

// ---------------------------------------------------------------------------------------------------
// DE problem configuration
// ---------------------------------------------------------------------------------------------------
extern "C" int dlra_rhs_set(dlra_handle h, const dlra_operator* A, const dlra_operator* B, const double* G, int64_t ldg,
                            const double* H, int64_t ldh, int q, const dlra_operator* D1, const dlra_operator* D2, double c_had) {
    DLRA_API_BEGIN(h)
    de_rhs_set(h, A, B, G, ldg, H, ldh, q, D1, D2, c_had);
    h->zcarry.valid = false;
    DLRA_API_END(h)
}

extern "C" int dlra_rhs_add_term(dlra_handle h, const dlra_operator* A, const dlra_operator* B) {
    DLRA_API_BEGIN(h)
    de_rhs_add_term(h, A, B);
    h->zcarry.valid = false;
    DLRA_API_END(h)
}

extern "C" int dlra_set_substepper(dlra_handle h, int flow, int ode, int nsub, double abstol, double reltol) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(flow >= 0 && flow <= 2, "flow must be DLRA_FLOW_K|S|L");
    DLRA_REQUIRE(ode >= DLRA_ODE_EULER && ode <= DLRA_ODE_TSIT5, "unknown sub-stepper");
    DLRA_REQUIRE(nsub >= 1, "nsub >= 1");
    SubStepperCfg c;
    c.ode = ode; c.nsub = nsub;
    if (abstol > 0) c.abstol = abstol;
    if (reltol > 0) c.reltol = reltol;
    c.maxiters = h->sub[flow].maxiters;
    h->sub[flow] = c;
    DLRA_API_END(h)
}

extern "C" int dlra_set_substepper_maxiters(dlra_handle h, int flow, int64_t maxiters) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(flow >= 0 && flow <= 2, "flow must be DLRA_FLOW_K|S|L");
    DLRA_REQUIRE(maxiters >= 1, "maxiters >= 1");
    h->sub[flow].maxiters = maxiters;
    DLRA_API_END(h)
}

// ---------------------------------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------------------------------
extern "C" int dlra_reconstruct(dlra_handle h, double* Y, int64_t ld) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(Y && ld >= h->n, "bad output");
    const int r = h->r;
    h->nscr.ensure(h->n * (int64_t)r, h->cx.stream);
    gemm_nn(h->cx, h->n, r, r, h->U, h->n, nullptr, 0, h->S, h->W, false, h->nscr.p, h->n, 1.0, 0.0);   // US
    gemm_nn(h->cx, h->n, r, (int)h->m, h->nscr.p, h->n, nullptr, 0, h->V, h->m, true, Y, ld, 1.0, 0.0); // US*V'
    DLRA_API_END(h)
}

extern "C" int dlra_reconstruct_error(dlra_handle h, const double* Yref, int64_t ld, double* rel_fro) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(Yref && ld >= h->n && rel_fro, "bad arguments");
    const int r = h->r;
    h->nscr.ensure(h->n * (int64_t)r, h->cx.stream);
    gemm_nn(h->cx, h->n, r, r, h->U, h->n, nullptr, 0, h->S, h->W, false, h->nscr.p, h->n, 1.0, 0.0);
    dim3 grid((unsigned)cdiv(h->n, 128), (unsigned)cdiv(h->m, 512));
    const int64_t nblocks = (int64_t)grid.x * grid.y;
    h->gws.ensure(2 * nblocks + 16, h->cx.stream);
    recon_err_kernel<<<grid, 128, (size_t)r * 8 * sizeof(double), h->cx.stream>>>(h->n, h->m, r, h->nscr.p, h->n, h->V, h->m, Yref, ld, h->gws.p);
    h->cx.launches++;
    DLRA_CUDA(cudaGetLastError());
    sum_pairs_kernel<<<1, 1024, 0, h->cx.stream>>>(nblocks, h->gws.p, h->scal_dev);
    h->cx.launches++;
    h->comm.allreduce_sum(h->scal_dev, 2, h->cx);
    double out[2];
    DLRA_CUDA(cudaMemcpyAsync(out, h->scal_dev, 16, cudaMemcpyDeviceToHost, h->cx.stream));
    DLRA_CUDA(cudaStreamSynchronize(h->cx.stream));
    *rel_fro = (out[1] > 0.0) ? sqrt(out[0] / out[1]) : sqrt(out[0]);
    DLRA_API_END(h)
}

// C+ = Q * diag(sigma_k > tol ? 1/sigma_k : 0) * P'   for C = P*diag(sigma)*Q'   (pinv(C, atol = tol), utils.jl:7)
__global__ void __launch_bounds__(256) pinv_from_svd_kernel(int r, const double* __restrict__ P, const double* __restrict__ sigma, const double* __restrict__ Q,
                                     int ld, double tol, double* __restrict__ Cp) {
    for (int e = threadIdx.x; e < r * r; e += blockDim.x) {
        const int i = e % r, j = e / r;
        double s = 0.0;
        for (int k = 0; k < r; ++k) {
            const double sg = sigma[k];
            if (sg > tol) s = fma(Q[i + (int64_t)k * ld] / sg, P[j + (int64_t)k * ld], s);
        }
        Cp[i + (int64_t)j * ld] = s;
    }
}

extern "C" int dlra_normal_component(dlra_handle h, const double* dY, int64_t ld, const double* C, int64_t ldc, double tol, double* out,
                                     int64_t ldo, double* fro_norm) {
    DLRA_API_BEGIN(h)
    settle_aux(h);
    DLRA_REQUIRE(dY && ld >= h->n, "bad dY");
    DLRA_REQUIRE(!out || ldo >= h->n, "bad output leading dimension");
    DLRA_REQUIRE(!C || ldc >= h->r, "bad C leading dimension");
    DLRA_REQUIRE(tol >= 0.0, "bad tolerance");
    Ctx& cx = h->cx;
    const int r = h->r, r2 = 2 * h->r;
    const int64_t n = h->n, m = h->m, W = h->W;
    h->kl_ready = false;                                   // the passes below reuse the pipelined step's partial buffers
    h->nscr.ensure(n * (int64_t)(3 * r), cx.stream);       // [ U | Kc | Kz ]
    h->mscr.ensure(m * (int64_t)r2, cx.stream);            // [ L | Z ]
    double *Xl = h->nscr.p, *Kc = Xl + n * r, *Kz = Xl + 2 * n * r;
    double *L = h->mscr.p, *Z = L + m * r;
    gemm_nn(cx, m, r, r, h->V, m, nullptr, 0, h->S, W, true, Z, m, 1.0, 0.0);                 // Z = V*S'
    fill_mat(cx, n, r, Kz, n, 0.0, 0.0);
    Delta d; d.A = dY; d.lda = ld;
    pass_KL(h, d, r, Z, m, h->U, n, Kz, n, L, m);                                             // Kz = dY*Z ; L = dY'*U (all-reduced)
    gram_mside(h, main_side(h), r, r, L, Z, h->T1);                                           // T1 = L'*Z = U'*dY*Z
    if (C) copy_mat(cx, r, r, C, ldc, false, h->Sh, W);
    else gram_mside(h, main_side(h), r, r, Z, Z, h->Sh);                                      // C = Z'*Z
    h->jws.ensure((int64_t)jacobi_ws_doubles(r), cx.stream);
    jacobi_svd(cx, r, h->Sh, (int)W, h->jws.p, h->Pm, (int)W, h->sig, h->Qm, (int)W, 0.0, r, nullptr, nullptr);
    pinv_from_svd_kernel<<<1, 256, 0, cx.stream>>>(r, h->Pm, h->sig, h->Qm, (int)W, tol, h->T2);
    cx.launches++;
    DLRA_CUDA(cudaGetLastError());
    gemm_nn(cx, n, r, r, h->U, n, nullptr, 0, h->T1, W, false, Kz, n, -1.0, 1.0);             // Kz -= U*(U'*dY*Z)
    gemm_nn(cx, n, r, r, Kz, n, nullptr, 0, h->T2, W, false, Kc, n, 1.0, 0.0);                // Kc = (I-UU')*dY*Z*C+
    copy_mat(cx, n, r, h->U, n, false, Xl, n);
    // N = dY − [U Kc]*[L Z]'
    if (fro_norm) {
        dim3 grid((unsigned)cdiv(n, 128), (unsigned)cdiv(m, 512));
        const int64_t nblocks = (int64_t)grid.x * grid.y;
        h->gws.ensure(2 * nblocks + 16, cx.stream);
        recon_err_kernel<<<grid, 128, (size_t)r2 * 8 * sizeof(double), cx.stream>>>(n, m, r2, Xl, n, L, m, dY, ld, h->gws.p);
        cx.launches++;
        DLRA_CUDA(cudaGetLastError());
        sum_pairs_kernel<<<1, 1024, 0, cx.stream>>>(nblocks, h->gws.p, h->scal_dev);
        cx.launches++;
        h->comm.allreduce_sum(h->scal_dev, 2, cx);
    }
    if (out) {
        DLRA_CUDA(cudaMemcpy2DAsync(out, (size_t)ldo * sizeof(double), dY, (size_t)ld * sizeof(double), (size_t)n * sizeof(double),
                                    (size_t)m, cudaMemcpyDeviceToDevice, cx.stream));
        gemm_nn(cx, n, r2, (int)m, Xl, n, nullptr, 0, L, m, true, out, ldo, -1.0, 1.0);
    }
    if (fro_norm) {
        double res[2];
        DLRA_CUDA(cudaMemcpyAsync(res, h->scal_dev, 16, cudaMemcpyDeviceToHost, cx.stream));
        DLRA_CUDA(cudaStreamSynchronize(cx.stream));
        *fro_norm = sqrt(res[0]);
    }
    DLRA_API_END(h)
}

extern "C" int dlra_set_profiling(dlra_handle h, int time_passes) {
    if (!h) return DLRA_EINVAL;
    h->time_passes = time_passes != 0;
    return DLRA_OK;
}

extern "C" int dlra_stats(dlra_handle h, int64_t* kernel_launches, int64_t* pass_launches, double* pass_ms_total,
                          double* pass_bytes_total, int reset) {
    DLRA_API_BEGIN(h)
    DLRA_CUDA(cudaStreamSynchronize(h->cx.stream));
    for (size_t i = 0; i < h->pass_events.size(); ++i) {
        auto& pr = h->pass_events[i];
        float ms = 0.f;
        DLRA_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        h->pass_ms += ms;
        h->kind_ms[h->pass_event_kind[i]] += ms;
        cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
    }
    h->pass_events.clear();
    h->pass_event_kind.clear();
    if (!h->phase_marks.empty()) {
        // marks named "step" open a step; every other mark closes a phase
        std::vector<std::pair<std::string, double>> acc;
        int steps = 0;
        for (size_t i = 0; i < h->phase_marks.size(); ++i) {
            const char* nm = h->phase_marks[i].first;
            if (strcmp(nm, "step") == 0) { ++steps; continue; }
            if (i == 0) continue;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->phase_marks[i - 1].second, h->phase_marks[i].second);
            bool found = false;
            for (auto& a : acc) if (a.first == nm) { a.second += ms; found = true; break; }
            if (!found) acc.emplace_back(nm, ms);
        }
        if (reset == 0 || getenv("DLRA_PHASES_ALWAYS")) {
            int dev = 0; cudaGetDevice(&dev);
            double tot = 0; for (auto& a : acc) tot += a.second;
            char line[2048];
            int off = snprintf(line, sizeof line, "[dlra phases] device %d rank %d/%d: %d steps, %.1f us/step on the main stream:", dev, h->comm.rank,
                               h->comm.nranks, steps, steps ? tot * 1e3 / steps : 0.0);
            for (auto& a : acc)
                if (off < (int)sizeof line - 64) off += snprintf(line + off, sizeof line - off, " %s %.1f", a.first.c_str(), steps ? a.second * 1e3 / steps : 0.0);
            snprintf(line + off, sizeof line - off, "\n");
            fputs(line, stderr);   // one write per rank: lines of different ranks do not interleave
        }
        for (auto& pm : h->phase_marks) cudaEventDestroy(pm.second);
        h->phase_marks.clear();
    }
    if (kernel_launches) *kernel_launches = h->cx.launches + h->ax.launches;
    if (pass_launches) *pass_launches = h->pass_launches;
    if (pass_ms_total) *pass_ms_total = h->pass_ms;
    if (pass_bytes_total) *pass_bytes_total = h->pass_bytes;
    if (reset) {
        h->cx.launches = 0; h->ax.launches = 0; h->pass_launches = 0; h->pass_ms = 0.0; h->pass_bytes = 0.0;
        for (int i = 0; i < 4; ++i) { h->kind_launches[i] = 0; h->kind_ms[i] = 0; h->kind_bytes[i] = 0; h->kind_flops[i] = 0; }
    }
    DLRA_API_END(h)
}

extern "C" int dlra_pass_breakdown(dlra_handle h, int64_t launches[4], double ms[4], double bytes[4], double flops[4]) {
    DLRA_API_BEGIN(h)
    int64_t a, b; double c, d;
    int rc = dlra_stats(h, &a, &b, &c, &d, 0);   // folds pending events into the per-kind sums
    DLRA_REQUIRE(rc == DLRA_OK, "stats failed");
    for (int i = 0; i < 4; ++i) {
        if (launches) launches[i] = h->kind_launches[i];
        if (ms) ms[i] = h->kind_ms[i];
        if (bytes) bytes[i] = h->kind_bytes[i];
        if (flops) flops[i] = h->kind_flops[i];
    }
    DLRA_API_END(h)
}

extern "C" int dlra_event_record(dlra_handle h, int slot) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(slot >= 0 && slot < 8, "event slot out of range");
    if (!h->user_events[slot]) DLRA_CUDA(cudaEventCreate(&h->user_events[slot]));
    DLRA_CUDA(cudaEventRecord(h->user_events[slot], h->cx.stream));
    DLRA_API_END(h)
}

extern "C" int dlra_event_elapsed_ms(dlra_handle h, int a, int b, double* ms) {
    DLRA_API_BEGIN(h)
    DLRA_REQUIRE(a >= 0 && a < 8 && b >= 0 && b < 8 && ms && h->user_events[a] && h->user_events[b], "bad event slots");
    DLRA_CUDA(cudaEventSynchronize(h->user_events[b]));
    float f = 0.f;
    DLRA_CUDA(cudaEventElapsedTime(&f, h->user_events[a], h->user_events[b]));
    *ms = f;
    DLRA_API_END(h)
}
