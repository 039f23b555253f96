// Collectives of the row-sharded multi-GPU path.  The reference is single-process and has no
// collective; here the partial sums of L / M / S are all-reduced and the TSQR R-factors all-gathered.  Two transports:
//
//  * P2P (default of the host mirror on one NVSwitch box): every rank exposes an exchange region through CUDA IPC; a collective is
//    "post my contribution + raise a sequence flag in every peer" followed by ONE kernel that waits for the flags and
//    reads the peers' contributions directly over NVLink (ld.relaxed.sys on mapped peer pointers), summing them in rank
//    order — bit-identical on every rank, ~10 us instead of NCCL's small-message latency, no extra copies.
//  * NCCL: libnccl.so.2 is dlopen'ed at dlra_comm_init time (the torch-bundled copy is reused when the host process
//    already loaded it), so libdlra.so itself has no link-time NCCL dependency and loads on CPU-only boxes.
#pragma once
#include "common.cuh"
#include <dlfcn.h>

namespace dlra {

constexpr int P2P_MAX_RANKS = 8;
constexpr int P2P_FLAG_STRIDE = 16;   // uint64 per flag slot (128 bytes: one line per writer rank)

struct P2PView {
    int nranks, rank;
    unsigned long long seq;
    unsigned long long* flags_local;            // this rank's flag block
    unsigned long long* flags_peer[P2P_MAX_RANKS];
    const double* data_peer[P2P_MAX_RANKS];     // peers' exchange buffers of the current parity (own included)
    double* data_local;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// ---- "LL" exchange (flag-in-data, the protocol NCCL uses for latency-bound messages): every double travels as ONE 16-byte
// store {lo32, flag, hi32, flag} written straight into slot [sender][element] of the RECEIVER's buffer over NVLink, and the
// receiver spins on its own memory until both flags carry the message's sequence number.  No fence, no separate flag
// round trip, no staging copy: a small all-reduce costs one NVLink write latency instead of ~50 us of
// fence + release/acquire handshakes (profiles/r02/multi_gpu_phases.txt).
struct LLView {
    int nranks, rank;
    unsigned int seq;            // never 0 (the buffers start zeroed)
    int fence;                   // 1: membar.sys after the pushes of a kernel (see ll_flush)
    int backoff;                 // > 0: nanoseconds to sleep after a failed poll (a tight volatile-load loop on the lines a peer is writing)
    int atomic_poll;             // 1: poll with a 128-bit compare-and-swap that never matches (served by the L2 slice that owns the line)
    int64_t cap;                 // elements per sender slot
    uint4* local;                // this rank's buffer of the current parity: [nranks][cap]
    uint4* peer[P2P_MAX_RANKS];  // the same buffer in every rank (own included)
};
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned int seq) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)b), "r"(seq), "r"((unsigned int)(b >> 32)), "r"(seq)
                 : "memory");
}
// 16-byte read performed AT the point of coherence: a compare-and-swap whose compare value cannot occur (the two flag words differ)
// and whose swap value equals it, so memory never changes.  Unlike a load it cannot be served from a copy of the line that some
// cache level took before the peer's store arrived.
__device__ __forceinline__ void ll_read_coherent(const uint4* p, unsigned int& lo, unsigned int& f1, unsigned int& hi, unsigned int& f2) {
    unsigned long long a, b;
    asm volatile(
        "{\n\t.reg .b128 d, c;\n\t"
        "mov.b128 c, {%3, %4};\n\t"
        "atom.global.sys.relaxed.cas.b128 d, [%2], c, c;\n\t"
        "mov.b128 {%0, %1}, d;\n\t}"
        : "=l"(a), "=l"(b) : "l"(p), "l"(0xdeadbeef00000000ULL), "l"(0xfeedface00000000ULL) : "memory");
    lo = (unsigned int)a; f1 = (unsigned int)(a >> 32); hi = (unsigned int)b; f2 = (unsigned int)(b >> 32);
}
__device__ __forceinline__ double ll_load(const uint4* p, unsigned int seq, int backoff = 0, int atomic_poll = 0) {
    unsigned int lo, f1, hi, f2;
    while (true) {
        if (atomic_poll) ll_read_coherent(p, lo, f1, hi, f2);
        else asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(p) : "memory");
        if (f1 == seq && f2 == seq) break;
        if (backoff > 0) __nanosleep(backoff);
    }
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
// After the last push of a kernel: nothing orders a posted remote store against the spin loop that follows it, so the hardware may
// keep the tail of the pushes in flight until the kernel retires.  With DLRA_LL_FENCE=1 every pushing thread drains its own stores
// (membar.sys: one NVLink round trip) before it starts to poll.  A/B in profiles/r02/multi_gpu_r02b.txt.
__device__ __forceinline__ void ll_flush(const LLView& v) {
    if (v.fence) __threadfence_system();
}
// one attempt (no spinning): many of these can be in flight before the first flag is inspected
__device__ __forceinline__ bool ll_try_load(const uint4* p, unsigned int seq, double& out, int atomic_poll = 0) {
    unsigned int lo, f1, hi, f2;
    if (atomic_poll) ll_read_coherent(p, lo, f1, hi, f2);
    else asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(p) : "memory");
    out = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    return f1 == seq && f2 == seq;
}
// element e of this rank's contribution -> every rank's buffer
__device__ __forceinline__ void ll_push(const LLView& v, int64_t e, double val) {
    for (int g = 0; g < v.nranks; ++g) ll_store(v.peer[g] + (size_t)v.rank * v.cap + e, val, v.seq);
}
// sum over ranks of element e, in rank order (bit-identical on every rank)
__device__ __forceinline__ double ll_sum(const LLView& v, int64_t e) {
    double s = 0.0;
    for (int g = 0; g < v.nranks; ++g) s += ll_load(v.local + (size_t)g * v.cap + e, v.seq, v.backoff, v.atomic_poll);
    return s;
}
// in-place all-reduce of a dense vector: one launch, every thread pushes its elements and then collects the peers' copies
__global__ void __launch_bounds__(256) ll_allreduce_kernel(LLView v, double* __restrict__ buf, int64_t count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) ll_push(v, e, buf[e]);
    ll_flush(v);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) buf[e] = ll_sum(v, e);
}
// dst[g*count + i] = rank g's src[i]
__global__ void __launch_bounds__(256) ll_allgather_kernel(LLView v, const double* __restrict__ src, double* __restrict__ dst, int64_t count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) ll_push(v, e, src[e]);
    ll_flush(v);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count * v.nranks; e += stride) {
        const int g = (int)(e / count);
        dst[e] = ll_load(v.local + (size_t)g * v.cap + (e - (int64_t)g * count), v.seq, v.backoff, v.atomic_poll);
    }
}

// copy `count` doubles into the local exchange buffer; the last CTA to finish raises this rank's flag in every peer
__global__ void __launch_bounds__(256) p2p_post_kernel(P2PView v, const double* __restrict__ src, int64_t count, unsigned int* ticket) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) v.data_local[i] = src[i];
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();   // one cumulative fence per CTA (the barrier ordered the other threads' stores before it)
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x < v.nranks) st_release_sys(v.flags_peer[threadIdx.x] + (size_t)v.rank * P2P_FLAG_STRIDE, v.seq);
    if (threadIdx.x == 0) *ticket = 0;
}

__device__ __forceinline__ void p2p_wait_all(const P2PView& v) {
    if (threadIdx.x < v.nranks) {
        const unsigned long long* f = v.flags_local + (size_t)threadIdx.x * P2P_FLAG_STRIDE;
        while (ld_acquire_sys(f) < v.seq) { }
    }
    __syncthreads();
}

// dst[i] = Σ_g peer_g[i] in rank order (identical result on every rank)
__global__ void __launch_bounds__(256) p2p_sum_kernel(P2PView v, double* __restrict__ dst, int64_t count) {
    p2p_wait_all(v);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int g = 0; g < v.nranks; ++g) s += ld_relaxed_sys(v.data_peer[g] + i);
        dst[i] = s;
    }
}
// dst[g*count + i] = peer_g[i]
__global__ void __launch_bounds__(256) p2p_gather_kernel(P2PView v, double* __restrict__ dst, int64_t count) {
    p2p_wait_all(v);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count * v.nranks; i += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(i / count);
        dst[i] = ld_relaxed_sys(v.data_peer[g] + (i - (int64_t)g * count));
    }
}

// Small-message all-reduce in ONE single-CTA launch: up to four little matrices (leading dimension ld) are packed into the
// exchange buffer, the rank's flag is raised in every peer, the peers' packs are summed in rank order and unpacked in place.
// Replaces (copy_mat x2, post, sum, copy_mat x2) of the M / core-increment reduction of a BUG step.
struct SmallMats {
    double* p[4]; int rows[4], cols[4]; int64_t ld[4]; int n;
};
__global__ void __launch_bounds__(256) p2p_small_allreduce_kernel(P2PView v, SmallMats sm) {
    int64_t off = 0;
    for (int q = 0; q < sm.n; ++q) {
        const int cnt = sm.rows[q] * sm.cols[q];
        for (int e = threadIdx.x; e < cnt; e += blockDim.x)
            v.data_local[off + e] = sm.p[q][(e % sm.rows[q]) + (int64_t)(e / sm.rows[q]) * sm.ld[q]];
        off += cnt;
    }
    __syncthreads();
    __threadfence_system();
    if (threadIdx.x < v.nranks) st_release_sys(v.flags_peer[threadIdx.x] + (size_t)v.rank * P2P_FLAG_STRIDE, v.seq);
    p2p_wait_all(v);
    off = 0;
    for (int q = 0; q < sm.n; ++q) {
        const int cnt = sm.rows[q] * sm.cols[q];
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            double s = 0.0;
            for (int g = 0; g < v.nranks; ++g) s += ld_relaxed_sys(v.data_peer[g] + off + e);
            sm.p[q][(e % sm.rows[q]) + (int64_t)(e / sm.rows[q]) * sm.ld[q]] = s;
        }
        off += cnt;
    }
}

__global__ void __launch_bounds__(256) ll_small_allreduce_kernel(LLView v, SmallMats sm) {
    int64_t off = 0;
    for (int q = 0; q < sm.n; ++q) {
        const int cnt = sm.rows[q] * sm.cols[q];
        for (int e = threadIdx.x; e < cnt; e += blockDim.x)
            ll_push(v, off + e, sm.p[q][(e % sm.rows[q]) + (int64_t)(e / sm.rows[q]) * sm.ld[q]]);
        off += cnt;
    }
    ll_flush(v);
    off = 0;
    for (int q = 0; q < sm.n; ++q) {
        const int cnt = sm.rows[q] * sm.cols[q];
        for (int e = threadIdx.x; e < cnt; e += blockDim.x)
            sm.p[q][(e % sm.rows[q]) + (int64_t)(e / sm.rows[q]) * sm.ld[q]] = ll_sum(v, off + e);
        off += cnt;
    }
}

// Row-sharded BUG tail in ONE single-CTA launch (r <= 16): all-reduce of M = U1ᵀU0 and of the core increment R = U1ᵀΔA·V1 (LL push /
// collect, rank order) followed by the core update  S <- M·S·Nᵀ + R  (unconventional.jl:154-155) on the reduced values.
__global__ void __launch_bounds__(256) ll_allreduce_core_kernel(LLView v, int r, double* __restrict__ M, double* __restrict__ R, int ld,
                                                               double* __restrict__ S, const double* __restrict__ Nm) {
    __shared__ double Ms[256], Rs[256], Ts[256], Ss[256];
    const int i = threadIdx.x % 16, j = threadIdx.x / 16;
    const bool in = i < r && j < r;
    if (in) {
        ll_push(v, threadIdx.x, M[i + (int64_t)j * ld]);
        ll_push(v, 256 + threadIdx.x, R[i + (int64_t)j * ld]);
    }
    ll_flush(v);
    const double ms = in ? ll_sum(v, threadIdx.x) : 0.0;
    const double rs = in ? ll_sum(v, 256 + threadIdx.x) : 0.0;
    if (in) { M[i + (int64_t)j * ld] = ms; R[i + (int64_t)j * ld] = rs; }
    Ms[threadIdx.x] = ms; Rs[threadIdx.x] = rs;
    Ss[threadIdx.x] = in ? S[i + (int64_t)j * ld] : 0.0;
    __syncthreads();
    {
        double t = 0.0;
        for (int l = 0; l < r; ++l) t = fma(Ms[i + 16 * l], Ss[l + 16 * j], t);
        Ts[threadIdx.x] = t;
    }
    __syncthreads();
    if (in) {
        double t = 0.0;
        for (int l = 0; l < r; ++l) t = fma(Ts[i + 16 * l], Nm[j + (int64_t)l * ld], t);
        S[i + (int64_t)j * ld] = t + Rs[threadIdx.x];
    }
}

struct Comm {
    struct UniqueId { char internal[128]; };
    int nranks = 1, rank = 0;
    // ---- NCCL transport
    void* lib = nullptr;
    void* comm = nullptr;  // ncclComm_t
    int (*pGetUniqueId)(void*) = nullptr;
    int (*pCommInitRank)(void**, int, /*ncclUniqueId by value*/ UniqueId, int) = nullptr;
    int (*pCommDestroy)(void*) = nullptr;
    int (*pAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*pAllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*pGetErrorString)(int) = nullptr;
    // ---- P2P transport
    bool p2p = false;
    char* xbuf = nullptr;             // local exchange region: [flags 8 x 128 B][data parity 0][data parity 1]
    size_t xdata_bytes = 0;           // bytes of ONE parity buffer
    char* xpeer[P2P_MAX_RANKS] = {};
    // two independent channels (own flags, parity buffers, sequence numbers, ticket): channel 0 serves the main stream, channel 1
    // the auxiliary stream (the L all-reduce runs there beside the K-side TSQR, whose R all-gather uses channel 0)
    static constexpr int NCHAN = 2;
    unsigned long long seq[NCHAN] = {0, 0};
    unsigned int* ticket = nullptr;   // device counters for the post kernels, one per channel
    // LL region (after the flag/parity regions): per channel [2 parity][P2P_MAX_RANKS senders][ll_cap] x 16 bytes
    int64_t ll_cap = 0;
    unsigned int ll_seq[NCHAN] = {0, 0};
    static constexpr int64_t LL_MAX_CAP = 262144;   // 64 MB per channel at most

    static constexpr size_t FLAG_BYTES = (size_t)P2P_MAX_RANKS * P2P_FLAG_STRIDE * sizeof(unsigned long long);

    static void* open_lib() {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) return h;
        }
        return nullptr;
    }
    void load() {
        if (lib) return;
        lib = open_lib();
        if (!lib) throw CudaError(3, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
        pGetUniqueId = (decltype(pGetUniqueId))dlsym(lib, "ncclGetUniqueId");
        pCommInitRank = (decltype(pCommInitRank))dlsym(lib, "ncclCommInitRank");
        pCommDestroy = (decltype(pCommDestroy))dlsym(lib, "ncclCommDestroy");
        pAllReduce = (decltype(pAllReduce))dlsym(lib, "ncclAllReduce");
        pAllGather = (decltype(pAllGather))dlsym(lib, "ncclAllGather");
        pGetErrorString = (decltype(pGetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!pGetUniqueId || !pCommInitRank || !pAllReduce || !pAllGather) throw CudaError(3, "libnccl.so.2 lacks required symbols");
    }
    void check(int rc, const char* what) {
        if (rc != 0) throw CudaError(3, std::string(what) + " failed: " + (pGetErrorString ? pGetErrorString(rc) : "nccl error"));
    }
    void init(int nranks_, int rank_, const void* id128) {
        load();
        UniqueId id;
        memcpy(id.internal, id128, 128);
        check(pCommInitRank(&comm, nranks_, id, rank_), "ncclCommInitRank");
        nranks = nranks_;
        rank = rank_;
        p2p = false;   // an explicit NCCL init selects the NCCL transport
    }

    // ---- P2P set-up -----------------------------------------------------------------------------------------------
    void p2p_alloc(size_t data_bytes) {
        if (xbuf) return;
        xdata_bytes = (data_bytes + 255) / 256 * 256;
        ll_cap = std::min<int64_t>((int64_t)(xdata_bytes / 8), LL_MAX_CAP);
        DLRA_CUDA(cudaMalloc(&xbuf, NCHAN * chan_bytes() + NCHAN * ll_chan_bytes()));
        DLRA_CUDA(cudaMemset(xbuf, 0, NCHAN * chan_bytes() + NCHAN * ll_chan_bytes()));
        DLRA_CUDA(cudaMalloc(&ticket, NCHAN * sizeof(unsigned int)));
        DLRA_CUDA(cudaMemset(ticket, 0, NCHAN * sizeof(unsigned int)));
    }
    size_t chan_bytes() const { return FLAG_BYTES + 2 * xdata_bytes; }
    size_t ll_chan_bytes() const { return (size_t)2 * P2P_MAX_RANKS * (size_t)ll_cap * 16; }
    bool ll_fits(int64_t count) const { return p2p && ll_cap > 0 && count <= ll_cap && !ll_disabled(); }
    static bool ll_disabled() { static const bool d = getenv("DLRA_NO_LL") != nullptr; return d; }
    LLView next_ll(int chan) {
        unsigned int sq = ++ll_seq[chan];
        if (sq == 0) sq = ++ll_seq[chan];   // 0 is the "empty" flag value
        const size_t par = (size_t)(sq & 1);
        const size_t base = NCHAN * chan_bytes() + (size_t)chan * ll_chan_bytes() + par * (size_t)P2P_MAX_RANKS * (size_t)ll_cap * 16;
        LLView v;
        static const int fence = (getenv("DLRA_LL_FENCE") && atoi(getenv("DLRA_LL_FENCE")) != 0) ? 1 : 0;
        static const int backoff = getenv("DLRA_LL_BACKOFF_NS") ? atoi(getenv("DLRA_LL_BACKOFF_NS")) : 0;
        v.nranks = nranks; v.rank = rank; v.seq = sq; v.cap = ll_cap; v.fence = fence; v.backoff = backoff;
        static const int atomic_poll = (getenv("DLRA_LL_ATOMIC_POLL") && atoi(getenv("DLRA_LL_ATOMIC_POLL")) != 0) ? 1 : 0;
        v.atomic_poll = atomic_poll;
        v.local = (uint4*)(xbuf + base);
        for (int g = 0; g < P2P_MAX_RANKS; ++g) v.peer[g] = (uint4*)((xpeer[g] ? xpeer[g] : xbuf) + base);
        return v;
    }
    void p2p_export(void* handle64) {
        cudaIpcMemHandle_t hd;
        DLRA_CUDA(cudaIpcGetMemHandle(&hd, xbuf));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        memcpy(handle64, &hd, 64);
    }
    void p2p_import(int nranks_, int rank_, const void* handles) {
        DLRA_REQUIRE(xbuf != nullptr, "dlra_p2p_export must be called first");
        DLRA_REQUIRE(nranks_ >= 1 && nranks_ <= P2P_MAX_RANKS && rank_ >= 0 && rank_ < nranks_, "bad P2P arguments");
        for (int g = 0; g < nranks_; ++g) {
            if (g == rank_) { xpeer[g] = xbuf; continue; }
            cudaIpcMemHandle_t hd;
            memcpy(&hd, (const char*)handles + (size_t)g * 64, 64);
            void* p = nullptr;
            DLRA_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
            xpeer[g] = (char*)p;
        }
        nranks = nranks_;
        rank = rank_;
        p2p = nranks_ > 1;
    }
    void destroy() {
        if (comm && pCommDestroy) pCommDestroy(comm);
        comm = nullptr;
        for (int g = 0; g < P2P_MAX_RANKS; ++g)
            if (xpeer[g] && xpeer[g] != xbuf) cudaIpcCloseMemHandle(xpeer[g]);
        if (xbuf) cudaFree(xbuf);
        if (ticket) cudaFree(ticket);
        xbuf = nullptr; ticket = nullptr; p2p = false;
    }
    // Buffer reuse is safe by stream order: a rank posts message s+2 of a channel (same parity buffer as s) only after its own
    // wait for message s+1, i.e. after every peer has posted s+1, which every peer does after it finished reading message s.
    P2PView next_view(int chan = 0) {
        const unsigned long long sq = ++seq[chan];
        const size_t par = (size_t)(sq & 1);
        const size_t cb = (size_t)chan * chan_bytes();
        P2PView v;
        v.nranks = nranks; v.rank = rank; v.seq = sq;
        v.flags_local = (unsigned long long*)(xbuf + cb);
        for (int g = 0; g < P2P_MAX_RANKS; ++g) {
            char* base = (xpeer[g] ? xpeer[g] : xbuf) + cb;
            v.flags_peer[g] = (unsigned long long*)base;
            v.data_peer[g] = (const double*)(base + FLAG_BYTES + par * xdata_bytes);
        }
        v.data_local = (double*)(xbuf + cb + FLAG_BYTES + par * xdata_bytes);
        return v;
    }
    unsigned int* ticket_of(int chan) { return ticket + chan; }

    // in-place sum over ranks of up to four small matrices: one launch on the P2P transport
    void allreduce_small_mats(SmallMats sm, double* staging, Ctx& cx) {
        if (nranks <= 1 || sm.n <= 0) return;
        int64_t total = 0;
        for (int q = 0; q < sm.n; ++q) total += (int64_t)sm.rows[q] * sm.cols[q];
        if (ll_fits(total) && total <= 16384) {
            ll_small_allreduce_kernel<<<1, 256, 0, cx.stream>>>(next_ll(0), sm);
            cx.launches++;
            DLRA_CUDA(cudaGetLastError());
            return;
        }
        if (p2p && (size_t)total * 8 <= xdata_bytes && total <= 65536) {
            P2PView v = next_view(0);
            p2p_small_allreduce_kernel<<<1, 256, 0, cx.stream>>>(v, sm);
            cx.launches++;
            DLRA_CUDA(cudaGetLastError());
            return;
        }
        // library transport: pack densely, one all-reduce, unpack
        int64_t off = 0;
        for (int q = 0; q < sm.n; ++q) {
            DLRA_CUDA(cudaMemcpy2DAsync(staging + off, (size_t)sm.rows[q] * 8, sm.p[q], (size_t)sm.ld[q] * 8, (size_t)sm.rows[q] * 8, sm.cols[q],
                                        cudaMemcpyDeviceToDevice, cx.stream));
            off += (int64_t)sm.rows[q] * sm.cols[q];
        }
        allreduce_sum(staging, total, cx);
        off = 0;
        for (int q = 0; q < sm.n; ++q) {
            DLRA_CUDA(cudaMemcpy2DAsync(sm.p[q], (size_t)sm.ld[q] * 8, staging + off, (size_t)sm.rows[q] * 8, (size_t)sm.rows[q] * 8, sm.cols[q],
                                        cudaMemcpyDeviceToDevice, cx.stream));
            off += (int64_t)sm.rows[q] * sm.cols[q];
        }
    }

    // in-place sum over ranks
    void allreduce_sum(double* buf, int64_t count, Ctx& cx) {
        if (nranks <= 1 || count <= 0) return;
        if (ll_fits(count)) {
            // every thread spins only for elements it pushed itself: no grid-wide dependency, any grid size is safe
            const int blocks = (int)std::min<int64_t>(cdiv(count, 256), 2 * (int64_t)cx.num_sms);
            ll_allreduce_kernel<<<blocks, 256, 0, cx.stream>>>(next_ll(0), buf, count);
            cx.launches++;
            DLRA_CUDA(cudaGetLastError());
            return;
        }
        if (p2p) {
            DLRA_REQUIRE((size_t)count * 8 <= xdata_bytes, "P2P exchange region too small for this message");
            P2PView v = next_view(0);
            const int blocks = (int)std::min<int64_t>(cdiv(count, 1024), cx.num_sms);
            p2p_post_kernel<<<blocks, 256, 0, cx.stream>>>(v, buf, count, ticket);
            p2p_sum_kernel<<<blocks, 256, 0, cx.stream>>>(v, buf, count);
            cx.launches += 2;
            DLRA_CUDA(cudaGetLastError());
            return;
        }
        check(pAllReduce(buf, buf, (size_t)count, 8 /*ncclDouble*/, 0 /*ncclSum*/, comm, cx.stream), "ncclAllReduce");
    }
    void allgather(const double* send, double* recv, int64_t count_per_rank, Ctx& cx) {
        if (nranks <= 1) {
            if (send != recv) DLRA_CUDA(cudaMemcpyAsync(recv, send, count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, cx.stream));
            return;
        }
        if (p2p) {
            DLRA_REQUIRE((size_t)count_per_rank * 8 <= xdata_bytes, "P2P exchange region too small for this message");
            P2PView v = next_view(0);
            const int blocks = (int)std::min<int64_t>(cdiv(count_per_rank * nranks, 1024), cx.num_sms);
            p2p_post_kernel<<<std::max(1, (int)std::min<int64_t>(cdiv(count_per_rank, 1024), cx.num_sms)), 256, 0, cx.stream>>>(v, send, count_per_rank, ticket);
            p2p_gather_kernel<<<blocks, 256, 0, cx.stream>>>(v, recv, count_per_rank);
            cx.launches += 2;
            DLRA_CUDA(cudaGetLastError());
            return;
        }
        check(pAllGather(send, recv, (size_t)count_per_rank, 8, comm, cx.stream), "ncclAllGather");
    }
};

}  // namespace dlra
