// Micro-benchmarks that fix the roofline denominators MEASURED_PEAKS.json lacks:
// fp64 DFMA peak, fp64 DMMA (mma.sync m8n8k4) peak, and streaming-read bandwidth
// through (a) LDG.128 and (b) cp.async.bulk (UBLKCP) + mbarrier rings.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Prints one JSON object; bench.py reads profiles/fp64_peaks_r01.json derived from it.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x * 1e-9;
    double x = 1.0000001, y = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) out[threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double seed) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = seed + threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[threadIdx.x] = s;
}

// DMMA fed from shared memory: per DMMA pair one A fragment (LDS.64) — the pass-kernel ratio at r=16.
__global__ void __launch_bounds__(256) dmma_lds_kernel(double* out, int iters, double seed) {
    __shared__ double sm[64 * 36];
    for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) sm[i] = seed + i * 1e-9;
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double b0 = 1.0 + lane * 1e-12, b1 = 1.0 - lane * 1e-12;
    const double* base = sm + (warp * 8 + (lane >> 2)) + (lane & 3) * 68;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double a = base[(k * 4 * 68 + it) & 1023];
            dmma884(c[2 * k][0], c[2 * k][1], a, b0);
            dmma884(c[2 * k + 1][0], c[2 * k + 1][1], a, b1);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) ldg_stream_kernel(const double2* __restrict__ in, size_t n2, double* out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double s = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 v0 = __ldg(in + i), v1 = __ldg(in + i + stride), v2 = __ldg(in + i + 2 * stride), v3 = __ldg(in + i + 3 * stride);
        s += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
    }
    for (; i < n2; i += stride) { double2 v = __ldg(in + i); s += v.x + v.y; }
    if (s == 123.456) out[0] = s;
}

// ---- bulk-copy ring: one producer thread issues cp.async.bulk of CHUNK bytes per column-like segment.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Each CTA streams its contiguous slice of `in`; stage = NSEG segments of SEGB bytes, segments strided by `seg_stride` bytes
// (mimics columns of a column-major tile). Consumers read every 16th double so smem isn't the limiter.
template <int STAGES>
__global__ void __launch_bounds__(288) bulk_stream_kernel(const char* __restrict__ in, size_t seg_stride, int nseg, int segb,
                                                          int tiles_per_cta, int ctas_per_panel, double* out) {
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    const int stage_bytes = nseg * segb;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t tile_stride = (size_t)segb;
    const char* base = in + (size_t)(blockIdx.x / ctas_per_panel) * nseg * seg_stride
                          + (size_t)(blockIdx.x % ctas_per_panel) * tiles_per_cta * tile_stride;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 8) {
        for (int t = 0; t < tiles_per_cta; ++t) {
            int s = t % STAGES; uint32_t ph = (t / STAGES) & 1;
            if (t >= STAGES) mbar_wait(&empty[s], ph ^ 1);
            if (lane == 0) mbar_expect_tx(&full[s], stage_bytes);
            __syncwarp();
            for (int g = lane; g < nseg; g += 32)
                bulk_g2s(smem + (size_t)s * stage_bytes + (size_t)g * segb, base + (size_t)t * tile_stride + (size_t)g * seg_stride, segb, &full[s]);
        }
    } else {
        double acc = 0;
        for (int t = 0; t < tiles_per_cta; ++t) {
            int s = t % STAGES; uint32_t ph = (t / STAGES) & 1;
            mbar_wait(&full[s], ph);
            const double* p = (const double*)(smem + (size_t)s * stage_bytes);
            for (int i = threadIdx.x; i < stage_bytes / 8; i += 256 * 16) acc += p[i];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 123.456) out[0] = acc;
    }
}

template <class F> static float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 1 << 20));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    {
        int iters = 20000;
        float ms = time_ms([&] { dfma_kernel<<<sms * 8, 256>>>(out, iters, 1.0); });
        double fl = 2.0 * 16 * iters * 256.0 * sms * 8;
        printf(", \"dfma_tflops\": %.2f", fl / ms * 1e-9);
        // sustained: ~2 s
        int reps = (int)(2000.0f / ms) + 1;
        float tot = time_ms([&] { for (int r = 0; r < reps; ++r) dfma_kernel<<<sms * 8, 256>>>(out, iters, 1.0); }, 1);
        printf(", \"dfma_tflops_sustained\": %.2f", fl * reps / tot * 1e-9);
    }
    {
        int iters = 20000;
        float ms = time_ms([&] { dmma_kernel<8><<<sms * 4, 256>>>(out, iters, 1.0); });
        double fl = 2.0 * 256 * 8 * iters * 8.0 * sms * 4;
        printf(", \"dmma884_tflops_8acc_8w_x4cta\": %.2f", fl / ms * 1e-9);
        ms = time_ms([&] { dmma_kernel<8><<<sms, 256>>>(out, iters, 1.0); });
        printf(", \"dmma884_tflops_8acc_8w_x1cta\": %.2f", fl / 4 / ms * 1e-9);
        ms = time_ms([&] { dmma_kernel<2><<<sms, 256>>>(out, iters, 1.0); });
        printf(", \"dmma884_tflops_2acc_8w_x1cta\": %.2f", fl / 16 / ms * 1e-9);
        ms = time_ms([&] { dmma_kernel<8><<<sms, 128>>>(out, iters, 1.0); });
        printf(", \"dmma884_tflops_8acc_4w_x1cta\": %.2f", fl / 8 / ms * 1e-9);
        ms = time_ms([&] { dmma_kernel<1><<<sms, 128>>>(out, iters, 1.0); });
        printf(", \"dmma884_tflops_1acc_4w_x1cta(latency)\": %.2f", fl / 64 / ms * 1e-9);
        int reps = 40;
        float one = time_ms([&] { dmma_kernel<8><<<sms * 4, 256>>>(out, iters, 1.0); });
        reps = (int)(2000.0f / one) + 1;
        float tot = time_ms([&] { for (int r = 0; r < reps; ++r) dmma_kernel<8><<<sms * 4, 256>>>(out, iters, 1.0); }, 1);
        printf(", \"dmma884_tflops_sustained\": %.2f", fl * reps / tot * 1e-9);
        ms = time_ms([&] { dmma_lds_kernel<<<sms * 2, 256>>>(out, iters, 1.0); });
        printf(", \"dmma884_lds_fed_tflops\": %.2f", 2.0 * 256 * 8 * iters * 8.0 * sms * 2 / ms * 1e-9);
    }
    {
        size_t bytes = (size_t)4 << 30;
        char* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
        for (int mult : {2, 4, 8}) {
            float ms = time_ms([&] { ldg_stream_kernel<<<sms * mult, 512>>>((const double2*)buf, bytes / 16, out); });
            printf(", \"ldg128_read_gbs_x%d\": %.1f", mult, bytes / ms * 1e-6);
        }
        // bulk ring: tile = 32 segments (columns) of SEGB bytes from a column-major matrix with ld = 65536 doubles.
        {
            const size_t ld_bytes = 65536 * 8;
            struct Cfg { int stages, nseg, segb; };
            for (Cfg c : {Cfg{4, 32, 1024}, Cfg{6, 32, 1024}, Cfg{4, 32, 512}, Cfg{8, 32, 512}, Cfg{3, 32, 2048}, Cfg{6, 64, 512}}) {
                int rows_tiles = (int)(ld_bytes / c.segb);   // tiles going down one column panel
                int cpp = 2;                                  // CTAs per column panel
                int tiles_per_cta = rows_tiles / cpp;
                int grid = sms;                               // panels used = grid / cpp, all distinct addresses
                size_t smem = (size_t)c.stages * c.nseg * c.segb;
                auto launch = [&](auto kern) {
                    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    float ms = time_ms([&] { kern<<<grid, 288, smem>>>(buf, ld_bytes, c.nseg, c.segb, tiles_per_cta, cpp, out); });
                    double moved = (double)grid * tiles_per_cta * c.nseg * c.segb;
                    printf(", \"bulk_read_gbs_st%d_seg%dx%dB\": %.1f", c.stages, c.nseg, c.segb, moved / ms * 1e-6);
                    printf(", \"bulk_mb_st%d_seg%dx%dB\": %.0f", c.stages, c.nseg, c.segb, moved / 1e6);
                };
                switch (c.stages) {
                    case 3: launch(bulk_stream_kernel<3>); break;
                    case 4: launch(bulk_stream_kernel<4>); break;
                    case 6: launch(bulk_stream_kernel<6>); break;
                    case 8: launch(bulk_stream_kernel<8>); break;
                }
            }
        }
        cudaFree(buf);
    }
    printf("}\n");
    return 0;
}
