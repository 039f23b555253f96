// Shared helpers for libdlra.so (sm_100a only).
#pragma once
#include <nvtx3/nvToolsExt.h>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <stdexcept>
#include <type_traits>

namespace dlra {

struct CudaError : std::runtime_error {
    int code;
    CudaError(int c, const std::string& s) : std::runtime_error(s), code(c) {}
};

#define DLRA_CUDA(expr)                                                                            \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            throw ::dlra::CudaError(2, b__);                                                       \
        }                                                                                          \
    } while (0)

#define DLRA_REQUIRE(cond, msg)                                                                    \
    do {                                                                                           \
        if (!(cond)) throw ::dlra::CudaError(1, std::string(msg) + " [" #cond "]");                \
    } while (0)

// NVTX range for the profilers' timelines (header-only NVTX v3: a no-op unless a tool injects itself)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// cudaFuncSetAttribute is per device: a host process that drives several GPUs (one handle each) must repeat it on each one.
// `mask` is a function-local static; returns true the first time the calling kernel wrapper runs on the current device.
static inline bool first_use_on_this_device(unsigned long long& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// launch context: engine stream + launch counter (dlra_stats)
struct Ctx {
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    int num_sms = 148;
    unsigned int* counters = nullptr;   // zero-initialised device counters for last-block-done reductions (self-resetting)
    // one-launch TSQR (tsqr_fused_kernel): [0] R factors published (running total), [1] epoch of the finished tree top, [2] time-outs
    unsigned int* sync = nullptr;
    unsigned int sync_arrivals = 0, sync_epoch = 0;   // host copies of the running totals (wrap-around arithmetic)
    // main stream only: tall products n x p times p x q go through the TMA/DMMA streaming kernels when they fit (pass_tma.cuh)
    bool (*tall_gemm)(void* eng, int64_t n, int p, int q, const double* A, int64_t lda, const double* B, int64_t ldb, bool transB,
                      double* C, int64_t ldc, double alpha, double beta) = nullptr;
    void* tall_eng = nullptr;
    // DLRA_PHASES=1: phase marks from helpers that only see the launch context (main stream only)
    void (*mark_fn)(void* eng, const char* name) = nullptr;
    void mark(const char* name) const { if (mark_fn) mark_fn(tall_eng, name); }
};

// compile-time loop: f(std::integral_constant<int, I>) for I = B .. E-1 (indices are constant expressions in the
// front end, so register arrays indexed by them are always promoted — nested `#pragma unroll` is not reliable for that)
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// ---------------------------------------------------------------------------------------------
// device-side primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// fp64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4 (the only native f64 MMA
// shape on sm_100a; m16n8k{4,8,16} lower to sequences of it).
// lane t: a = A[t/4][t%4], b = B[t%4][t/4], c0/c1 = C[t/4][2*(t%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / TMA (cp.async.bulk.tensor) wrappers ------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
// 2-D tiled TMA load: coordinates (c0 = fastest dim = row index, c1 = column index). SASS: UTMALDG.2D
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset in every CTA of `mask` and each destination's mbarrier
// (same offset) receives the complete_tx for the bytes written there.
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5, %6;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask), "l"(policy)
        : "memory");
}
// ---- thread-block clusters -------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, uint32_t rank) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}" ::"r"(smem_u32(b)), "r"(rank)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace dlra
