// Fused streaming contraction over the dense increment ΔA (n x m, column-major) for sm_100a:
//     K[rows, c] += Σ_j ΔA[rows, j]·Vf[j, c]          (K-use, complete per row panel)
//     Lpart[cta][j, c] = Σ_{rows of the CTA's panels} ΔA[rows, j]·Uf[rows, c]   (L-use, per-CTA running partial, reduced afterwards)
// in ONE read of ΔA.  ΔA = A − Aprev is formed in shared memory when a previous snapshot is given
// (`Δy .= ycurr - yprev`, projector_splitting.jl:119-121), so the increment never exists in HBM.
//
// Mapping to the hardware
//  * one persistent CTA per SM owns row panels of 64·nsub rows and sweeps all column tiles (32 columns);
//    a producer warp streams 64 x 32 stages through a ring with TMA (cp.async.bulk.tensor.2d, SASS UTMALDG)
//    + mbarrier full/empty pairs; ΔA tiles carry an L2 evict_first hint, the factor tiles evict_last.
//  * tiles land as 16-row boxes in SWIZZLE_128B layout; the MMA row blocks use the row permutation
//    {0,1,8,9,2,3,10,11}+4b so that BOTH fragment patterns (8 rows x 4 cols for the K-use, 4 rows x 8 cols for the
//    L-use) read shared memory bank-conflict free.
//  * math runs on the fp64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 — tcgen05/UMMA has no f64 kind).
//    Each of the 8 consumer warps owns 8 rows of every stage for both uses; K accumulators stay in registers
//    for the whole panel, L accumulators for one column tile; a reducer warp of the service warpgroup sums the 8 warps'
//    tiles (mbarrier hand-off, no CTA barrier) into one running partial per CTA, and l_finalize_kernel reduces the
//    per-CTA partials in fixed order (deterministic, no atomics).
//  * r is processed in chunks of RT in {8, 16} factor columns: at RT = 16 the pass is already fp64-bound
//    (AI = r/2 flop/B vs a ridge of ~5.6), so wider ranks re-stream ΔA per chunk at no cost in time.
#pragma once
#include "engine.cuh"

namespace dlra {

constexpr int PT_SI = 64;        // rows per stage
constexpr int PT_TJ = 32;        // columns per stage
constexpr int PT_NSUB_MAX = 7;   // stages (row sub-tiles) per panel held in K accumulators
constexpr int PT_CONSUMERS = 8;  // consumer warps
constexpr int PT_THREADS = (PT_CONSUMERS + 4) * 32;   // + one service warpgroup: TMA producer warp + three L-reducer warps
constexpr int PT_NRED = 3;       // reducer warps (one warp alone is starved by the consumers' DMMA stream: its DADDs share the fp64 pipe)
constexpr int PT_MAX_CLUSTER = 4;  // CTAs of a cluster split the factor columns (16 each) and share every ΔA tile through TMA multicast
constexpr int PT_LRED_LD = PT_TJ + 2;                 // padded column stride of the L reduction buffer (conflict-free)
constexpr int PT_TBYTES = PT_SI * PT_TJ * 8;  // 16 KB
constexpr int PT_BOXBYTES = 16 * PT_TJ * 8;   // 4 KB: one 16-row box

struct PassParams {
    int64_t n, m;
    int rc;          // valid factor columns in this chunk (<= RT)
    int nsub;        // row sub-tiles per panel
    int npanels;
    int ntj;         // column tiles
    double* K; int64_t ldk;       // K-use output (+=, or = when kstore), may be null
    int kstore;
    double* Lpart; int64_t ldlp;  // [gridDim.x][ldlp x RT] per-CTA partial L (ldlp >= m), may be null
};

template <int RT, bool DO_K, bool DO_L, bool DIFF>
struct PassSmem {
    // row sub-tiles per panel: bounded by the K accumulators a consumer thread can hold next to the L accumulators
    static constexpr int NSUBM = (RT == 32 && DO_L) ? 3 : PT_NSUB_MAX;
    static constexpr int VBYTES = DO_K ? PT_TJ * RT * 8 : 0;
    static constexpr int STAGE_BYTES = ((PT_TBYTES * (DIFF ? 2 : 1) + VBYTES + 1023) / 1024) * 1024;
    static constexpr int UBOX_BYTES = 16 * RT * 8;
    static constexpr int UPANEL_BYTES = DO_L ? NSUBM * 4 * UBOX_BYTES : 0;
    static constexpr int LRED_BYTES = DO_L ? ((PT_CONSUMERS * PT_LRED_LD * RT * 8 + 1023) / 1024) * 1024 : 0;
    static constexpr int BUDGET = 225 * 1024;
    static constexpr int NST_RAW = (BUDGET - UPANEL_BYTES - LRED_BYTES - 1024) / STAGE_BYTES;
    static constexpr int NST = NST_RAW > 8 ? 8 : NST_RAW;
    static constexpr int TOTAL = NST * STAGE_BYTES + UPANEL_BYTES + LRED_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int RT, bool DO_K, bool DO_L, bool DIFF>
__global__ void __launch_bounds__(PT_THREADS, 1)
pass_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapP,
            const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapV, const PassParams prm) {
    using SM = PassSmem<RT, DO_K, DO_L, DIFF>;
    constexpr int NST = SM::NST;
    constexpr int NB = RT / 8;
    constexpr int LD = PT_LRED_LD;
    constexpr int NSUBM = SM::NSUBM;
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte aligned carve-up (SWIZZLE_128B boxes need it)
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* stages = base;
    unsigned char* upanel = stages + (size_t)NST * SM::STAGE_BYTES;
    double* lred = reinterpret_cast<double*>(upanel + SM::UPANEL_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(lred) + SM::LRED_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + NST;
    uint64_t* panel_done = bars + 2 * NST;
    uint64_t* lfull = bars + 2 * NST + 1;
    uint64_t* lfree = bars + 2 * NST + 2;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Thread-block cluster (1..4 CTAs): CTA q owns factor columns [RT*q, RT*q + RT) — its own V tile, U panel, K columns and
    // L partial — while the ΔA stage is fetched ONCE per cluster: CTA q issues boxes b with b % cl_size == q as multicast
    // loads that land in every CTA of the cluster.  A stage slot may be refilled only when the consumers of ALL CTAs have
    // released it, so every consumer warp arrives on the empty barrier of every CTA of the cluster.
    const uint32_t cl_size = cluster_nctarank(), cl_rank = cluster_ctarank();
    const int coff = (int)cl_rank * RT;
    const int nclusters = (int)(gridDim.x / cl_size), cluster_id = (int)(blockIdx.x / cl_size);
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], PT_CONSUMERS * cl_size); }
        mbar_init(panel_done, PT_CONSUMERS);
        mbar_init(lfull, PT_CONSUMERS);
        mbar_init(lfree, PT_NRED);
        mbar_fence_init();
    }
    __syncthreads();
    if (cl_size > 1) cluster_sync_all();   // every CTA's barriers exist before the first remote arrive / multicast

    const int nsub = prm.nsub;
    // register re-partitioning (sm_90+ setmaxnreg): the service warpgroup gives its registers to the two consumer warpgroups
    if (warp >= PT_CONSUMERS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (warp == PT_CONSUMERS) {
        // ------------------------------ producer: one elected lane drives the TMA engine -------------------
        if (lane == 0) {
            prefetch_tmap(&mapA);
            if (DIFF) prefetch_tmap(&mapP);
            if (DO_L) prefetch_tmap(&mapU);
            if (DO_K) prefetch_tmap(&mapV);
            const uint64_t pol_stream = policy_evict_first();
            const uint64_t pol_keep = policy_evict_last();
            const uint16_t mc_mask = (uint16_t)((1u << cl_size) - 1u);
            int stage = 0; uint32_t phase = 0; uint32_t pd_phase = 0; bool first_panel = true;
            for (int panel = cluster_id; panel < prm.npanels; panel += nclusters) {
                if (!first_panel) { mbar_wait(panel_done, pd_phase); pd_phase ^= 1; }  // U panel region is free again
                first_panel = false;
                const int row0 = panel * nsub * PT_SI;
                for (int jt = 0; jt < prm.ntj; ++jt) {
                    for (int s = 0; s < nsub; ++s) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint32_t bytes = PT_TBYTES * (DIFF ? 2 : 1);
                        if (DO_K && s == 0) bytes += SM::VBYTES;
                        if (DO_L && jt == 0) bytes += 4 * SM::UBOX_BYTES;
                        mbar_expect_tx(&full[stage], bytes);
                        unsigned char* sb = stages + (size_t)stage * SM::STAGE_BYTES;
                        const int r = row0 + s * PT_SI;
                        if (cl_size == 1) {
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                tma_load_2d(sb + b * PT_BOXBYTES, &mapA, r + 16 * b, jt * PT_TJ, &full[stage], pol_stream);
                                if (DIFF) tma_load_2d(sb + PT_TBYTES + b * PT_BOXBYTES, &mapP, r + 16 * b, jt * PT_TJ, &full[stage], pol_stream);
                            }
                        } else {
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                if ((uint32_t)b % cl_size != cl_rank) continue;   // the peers fetch the other boxes for everybody
                                tma_load_2d_mc(sb + b * PT_BOXBYTES, &mapA, r + 16 * b, jt * PT_TJ, &full[stage], mc_mask, pol_stream);
                                if (DIFF) tma_load_2d_mc(sb + PT_TBYTES + b * PT_BOXBYTES, &mapP, r + 16 * b, jt * PT_TJ, &full[stage], mc_mask, pol_stream);
                            }
                        }
                        if (DO_K && s == 0) {
#pragma unroll
                            for (int b = 0; b < 2; ++b)
                                tma_load_2d(sb + PT_TBYTES * (DIFF ? 2 : 1) + b * (16 * RT * 8), &mapV, jt * PT_TJ + 16 * b, coff, &full[stage], pol_keep);
                        }
                        if (DO_L && jt == 0) {
#pragma unroll
                            for (int b = 0; b < 4; ++b)
                                tma_load_2d(upanel + (size_t)(s * 4 + b) * SM::UBOX_BYTES, &mapU, r + 16 * b, coff, &full[stage], pol_keep);
                        }
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (DO_L) {
        // ------------------------------ L reducers: three warps sum the 8 consumer warps' partial L tiles off the critical
        // path; warp wr owns the factor columns i = wr, wr + 3, ... of the tile
        const int wr = warp - PT_CONSUMERS - 1;
        constexpr int NI = (RT + PT_NRED - 1) / PT_NRED;
        uint32_t ph = 0;
        // one partial per CTA: the first panel stores, later panels of this CTA accumulate (fixed order => deterministic)
        double* lp = prm.Lpart + (size_t)blockIdx.x * prm.ldlp * RT;
        for (int panel = cluster_id; panel < prm.npanels; panel += nclusters) {
            const bool first = (panel == cluster_id);
            for (int jt = 0; jt < prm.ntj; ++jt) {
                const int64_t col = (int64_t)jt * PT_TJ + lane;
                const bool okc = col < prm.m;
                double prev[NI];
                // the previous panels' running sum is fetched BEFORE waiting for the consumers (latency fully hidden)
#pragma unroll
                for (int q = 0; q < NI; ++q) {
                    const int i = wr + q * PT_NRED;
                    prev[q] = (!first && okc && i < RT) ? __ldcg(lp + col + (int64_t)i * prm.ldlp) : 0.0;
                }
                mbar_wait(lfull, ph);
#pragma unroll
                for (int q = 0; q < NI; ++q) {   // element (c = i, j = lane)
                    const int i = wr + q * PT_NRED;
                    if (i < RT) {
                        double sum = prev[q];
#pragma unroll
                        for (int w = 0; w < PT_CONSUMERS; ++w) sum += lred[(size_t)w * LD * RT + i * LD + lane];
                        if (okc) __stcg(lp + col + (int64_t)i * prm.ldlp, sum);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(lfree);
                ph ^= 1;
            }
        }
    }
    } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");

    // ---------------------------------- consumers: 8 warps, DMMA ------------------------------------------
    const int g = lane >> 2, k = lane & 3;
    const int wbox = warp >> 1, wblk = warp & 1;
    // K-use fragment (8 rows x 4 cols): row = prow(wblk, g), col = 4*ks + k
    const int prow = (g & 1) + ((g >> 1) & 1) * 8 + (g >> 2) * 2 + wblk * 4;
    const uint32_t offKe = k * 128 + (((prow >> 1) ^ k) << 4) + (prow & 1) * 8;          // even ks
    const uint32_t offKo = k * 128 + (((prow >> 1) ^ (4 + k)) << 4) + (prow & 1) * 8;    // odd ks
    // L-use fragment (4 rows x 8 cols): row = lrow(kk, k), col = 8*cb + g
    uint32_t offL[2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const int lrow = (k & 1) + ((k >> 1) & 1) * 8 + 2 * kk + 4 * wblk;
        offL[kk] = g * 128 + (((lrow >> 1) ^ g) << 4) + (lrow & 1) * 8;
    }
    // V fragment of the K-use: Vf[4*ks + k][8*nb + g] from two swizzled 16-row boxes ([c][16 rows], 128 B per column)
    uint32_t offV[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) offV[q] = g * 128 + ((((2 * q) + (k >> 1)) ^ g) << 4) + (k & 1) * 8;
    // rows of this thread's C fragments inside a stage
    const int crow = wbox * 16 + prow;

    int stage = 0; uint32_t phase = 0; uint32_t lfree_ph = 1;   // first wait on a fresh barrier passes
    for (int panel = cluster_id; panel < prm.npanels; panel += nclusters) {
        const int64_t row0 = (int64_t)panel * nsub * PT_SI;
        double kacc[NSUBM][NB][2];
        if (DO_K) {
#pragma unroll
            for (int s = 0; s < NSUBM; ++s)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) { kacc[s][nb][0] = 0.0; kacc[s][nb][1] = 0.0; }
        }
        for (int jt = 0; jt < prm.ntj; ++jt) {
            double lacc[4][NB][2];
            double vf[8][NB];
            if (DO_L) {
#pragma unroll
                for (int cb = 0; cb < 4; ++cb)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) { lacc[cb][nb][0] = 0.0; lacc[cb][nb][1] = 0.0; }
            }
#pragma unroll
            for (int s = 0; s < NSUBM; ++s) {
                if (s < nsub) {
                    mbar_wait(&full[stage], phase);
                    const unsigned char* sb = stages + (size_t)stage * SM::STAGE_BYTES;
                    const unsigned char* tb = sb + wbox * PT_BOXBYTES;
                    if (DO_K && s == 0) {
                        const unsigned char* vt = sb + PT_TBYTES * (DIFF ? 2 : 1);
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb)
                                vf[ks][nb] = *reinterpret_cast<const double*>(vt + (ks >> 2) * (16 * RT * 8) + nb * 1024 + offV[ks & 3]);
                    }
                    if (DO_K) {
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            const uint32_t off = ((ks & 1) ? offKo : offKe) + ks * 512;
                            double a = *reinterpret_cast<const double*>(tb + off);
                            if (DIFF) a -= *reinterpret_cast<const double*>(tb + PT_TBYTES + off);   // ΔA = A − Aprev
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) dmma884(kacc[s][nb][0], kacc[s][nb][1], a, vf[ks][nb]);
                        }
                    }
                    if (DO_L) {
                        const unsigned char* ub = upanel + (size_t)(s * 4 + wbox) * SM::UBOX_BYTES;
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            double uf[NB];
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) uf[nb] = *reinterpret_cast<const double*>(ub + offL[kk] + nb * 1024);
#pragma unroll
                            for (int cb = 0; cb < 4; ++cb) {
                                const uint32_t off = offL[kk] + cb * 1024;
                                double a = *reinterpret_cast<const double*>(tb + off);
                                if (DIFF) a -= *reinterpret_cast<const double*>(tb + PT_TBYTES + off);
#pragma unroll
                                for (int nb = 0; nb < NB; ++nb) dmma884(lacc[cb][nb][0], lacc[cb][nb][1], a, uf[nb]);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (cl_size == 1) mbar_arrive(&empty[stage]);
                        else for (uint32_t q = 0; q < cl_size; ++q) mbar_arrive_remote(&empty[stage], q);
                    }
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
            if (DO_L) {
                // hand this warp's partial L tile (32 cols x RT) to the reducer warps
                mbar_wait(lfree, lfree_ph);
                lfree_ph ^= 1;
                double* mine = lred + (size_t)warp * LD * RT;
#pragma unroll
                for (int cb = 0; cb < 4; ++cb)
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        mine[(8 * nb + 2 * k) * LD + 8 * cb + g] = lacc[cb][nb][0];
                        mine[(8 * nb + 2 * k + 1) * LD + 8 * cb + g] = lacc[cb][nb][1];
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(lfull);
            }
        }
        if (DO_K) {
#pragma unroll
            for (int s = 0; s < NSUBM; ++s) {
                if (s < nsub) {
                    const int64_t row = row0 + s * PT_SI + crow;
                    if (row < prm.n) {
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = coff + 8 * nb + 2 * k + e;
                                if (c < prm.rc) {
                                    double* dst = prm.K + row + (int64_t)c * prm.ldk;
                                    *dst = prm.kstore ? kacc[s][nb][e] : *dst + kacc[s][nb][e];
                                }
                            }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(panel_done);
    }
    }
    // no CTA may exit while a peer can still multicast into its shared memory or arrive on its barriers
    if (cl_size > 1) { __syncwarp(); cluster_sync_all(); }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        DLRA_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) throw CudaError(2, "cuTensorMapEncodeTiled not available from the driver");
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

inline CUtensorMap make_map_2d(const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols, bool swizzle128) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw CudaError(2, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)rc) + ")");
    return m;
}

inline bool tma_ok(const double* p, int64_t ld) { return p && (((uintptr_t)p) & 15) == 0 && (ld % 2) == 0; }

inline bool tma_pass_supported(int64_t n, int64_t m, const Delta& d) {
    if (n < 64 || m < 8 || (n % 2) != 0 || (m % 2) != 0) return false;  // TMA strides must be multiples of 16 bytes
    if (n >= ((int64_t)1 << 31) || m >= ((int64_t)1 << 31)) return false;
    if (!tma_ok(d.A, d.lda)) return false;
    if (d.Aprev && !tma_ok(d.Aprev, d.ldap)) return false;
    return true;
}

// rows per panel = 64·nsub; `workers` = clusters (or CTAs) that sweep panels concurrently
inline int choose_nsub(int64_t n, int workers, int nsub_max = PT_NSUB_MAX) {
    const int64_t subtiles = cdiv(n, PT_SI);
    int best = 1; double best_cost = 1e300;
    for (int ns = 1; ns <= nsub_max; ++ns) {
        const int64_t panels = cdiv(subtiles, ns);
        const double cost = (double)cdiv(panels, workers) * ns + 0.02 * (double)panels / workers;  // makespan + partial-L overhead
        if (cost < best_cost - 1e-12) { best_cost = cost; best = ns; }
    }
    return best;
}

// co-resident clusters of `csize` CTAs for a kernel (depends on the GPC shapes of the part; queried once per device and size)
template <class Kern>
inline int max_active_clusters(Kern kern, int csize, int smem_bytes, int num_sms) {
    if (csize <= 1) return num_sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(csize * (num_sms / csize)));
    cfg.blockDim = dim3(PT_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    DLRA_CUDA(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
    DLRA_REQUIRE(ncl >= 1, "the device cannot co-schedule a cluster of this size");
    return std::min(ncl, num_sms / csize);
}

// shape of one pass launch: `csize` CTAs per cluster (1 = plain launch), nclusters clusters sweeping the row panels
struct PassGrid {
    int csize = 1, nclusters = 1, nsub = 1, npanels = 1;
    int ctas() const { return csize * nclusters; }
};

template <int RT, bool DO_K, bool DO_L, bool DIFF>
inline PassGrid pass_grid(dlra_engine* e, int csize) {
    using SM = PassSmem<RT, DO_K, DO_L, DIFF>;
    auto kern = pass_kernel<RT, DO_K, DO_L, DIFF>;
    static unsigned long long attr_devs = 0;
    static int cached[64][PT_MAX_CLUSTER + 1] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (first_use_on_this_device(attr_devs))
        DLRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    DLRA_REQUIRE(csize >= 1 && csize <= PT_MAX_CLUSTER, "cluster size out of range");
    int& mc = cached[dev & 63][csize];
    if (mc == 0) {
        mc = max_active_clusters(kern, csize, SM::TOTAL, e->cx.num_sms);
        if (getenv("DLRA_DEBUG")) fprintf(stderr, "[dlra] pass_kernel<%d,%d,%d,%d> cluster %d: %d co-resident clusters (%d of %d SMs), %d stages, %d B smem\n",
                                          RT, (int)DO_K, (int)DO_L, (int)DIFF, csize, mc, mc * csize, e->cx.num_sms, SM::NST, SM::TOTAL);
    }
    PassGrid g;
    g.csize = csize;
    g.nsub = choose_nsub(e->n, mc, SM::NSUBM);
    g.npanels = (int)cdiv(cdiv(e->n, PT_SI), g.nsub);
    g.nclusters = std::min(g.npanels, mc);
    return g;
}

// one launch: factor columns [0, rc) of Vf / Uf / K, RT per CTA, g.csize >= ceil(rc / RT) CTAs per cluster.
// Lpart: g.ctas() partials of ldlp x RT doubles; the partials of the factor columns [RT*q, RT*q + RT) are those of the CTAs
// with rank q in their cluster, i.e. blocks q, q + csize, q + 2*csize, ...
// mcols > 0: the streamed matrix has mcols columns instead of e->m (a tall factor in the role of ΔA: tall_gemm below);
// kstore: K = ΔA·Vf instead of K += ; timed = false keeps such launches out of the pass statistics
struct PassOpts {
    int64_t mcols = 0;
    bool kstore = false;
    bool timed = true;
};

template <int RT, bool DO_K, bool DO_L, bool DIFF>
inline void launch_pass(dlra_engine* e, const Delta& d, int rc, const double* Vf, int64_t ldv, const double* Uf, int64_t ldu,
                        double* K, int64_t ldk, double* Lpart, int64_t ldlp, const PassGrid& g, const PassOpts& po = PassOpts()) {
    const int64_t mm = po.mcols > 0 ? po.mcols : e->m;
    using SM = PassSmem<RT, DO_K, DO_L, DIFF>;
    static_assert(SM::NST >= 2, "pipeline needs at least two stages");
    auto kern = pass_kernel<RT, DO_K, DO_L, DIFF>;
    CUtensorMap mapA = make_map_2d(d.A, e->n, mm, d.lda, 16, PT_TJ, true);
    CUtensorMap mapP = DIFF ? make_map_2d(d.Aprev, e->n, mm, d.ldap, 16, PT_TJ, true) : mapA;
    CUtensorMap mapU = DO_L ? make_map_2d(Uf, e->n, rc, ldu, 16, RT, true) : mapA;
    CUtensorMap mapV = DO_K ? make_map_2d(Vf, mm, rc, ldv, 16, RT, true) : mapA;
    PassParams prm;
    prm.n = e->n; prm.m = mm; prm.rc = rc; prm.nsub = g.nsub; prm.npanels = g.npanels; prm.ntj = (int)cdiv(mm, PT_TJ);
    prm.K = K; prm.ldk = ldk; prm.Lpart = Lpart; prm.ldlp = ldlp; prm.kstore = po.kstore ? 1 : 0;
    const double bytes = (double)e->n * (double)mm * 8.0 * (DIFF ? 2.0 : 1.0);
    const double flops = 2.0 * (double)e->n * (double)mm * rc * ((DO_K ? 1 : 0) + (DO_L ? 1 : 0));
    if (po.timed) pass_timer_begin(e, bytes, (DO_K && DO_L) ? 0 : (DO_K ? 1 : 2), flops);
    if (g.csize == 1) {
        kern<<<g.ctas(), PT_THREADS, SM::TOTAL, e->cx.stream>>>(mapA, mapP, mapU, mapV, prm);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)g.ctas());
        cfg.blockDim = dim3(PT_THREADS);
        cfg.dynamicSmemBytes = SM::TOTAL;
        cfg.stream = e->cx.stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)g.csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        DLRA_CUDA(cudaLaunchKernelEx(&cfg, kern, mapA, mapP, mapU, mapV, prm));
    }
    if (po.timed) pass_timer_end(e);
    e->cx.launches++;
    DLRA_CUDA(cudaGetLastError());
}

// runtime (K?, L?, DIFF?) -> template instance
template <int RT>
inline PassGrid pass_grid_rt(dlra_engine* e, bool dk, bool dl, bool df, int csize) {
#define DLRA_PASS_CASE(a, b, c) if (dk == a && dl == b && df == c) return pass_grid<RT, a, b, c>(e, csize);
    DLRA_PASS_CASE(true, true, false)
    DLRA_PASS_CASE(true, true, true)
    DLRA_PASS_CASE(true, false, false)
    DLRA_PASS_CASE(true, false, true)
    DLRA_PASS_CASE(false, true, false)
    DLRA_PASS_CASE(false, true, true)
#undef DLRA_PASS_CASE
    throw CudaError(1, "pass without outputs");
}
template <int RT>
inline void launch_pass_rt(dlra_engine* e, const Delta& d, int rc, const double* Vf, int64_t ldv, const double* Uf, int64_t ldu,
                           double* K, int64_t ldk, double* Lpart, int64_t ldlp, const PassGrid& g, const PassOpts& po = PassOpts()) {
    const bool dk = K != nullptr, dl = Lpart != nullptr, df = d.Aprev != nullptr;
#define DLRA_PASS_CASE(a, b, c) \
    if (dk == a && dl == b && df == c) return launch_pass<RT, a, b, c>(e, d, rc, Vf, ldv, Uf, ldu, K, ldk, Lpart, ldlp, g, po);
    DLRA_PASS_CASE(true, true, false)
    DLRA_PASS_CASE(true, true, true)
    DLRA_PASS_CASE(true, false, false)
    DLRA_PASS_CASE(true, false, true)
    DLRA_PASS_CASE(false, true, false)
    DLRA_PASS_CASE(false, true, true)
#undef DLRA_PASS_CASE
}

// Fused tail of the L-use: fixed-order sum of the per-CTA partials, cross-rank sum over NVLink peer memory (P2P transport,
// row-sharded runs) and the initial term  L += Vi·Siᵀ  (the reference's `mul!(VS, V, S')` before the L-step,
// unconventional.jl:145) in ONE launch.  XR: post the local sum to the exchange buffer, the last CTA raises this rank's
// sequence flag in every peer, then all CTAs wait for the peers' flags and add the peers' slices in rank order.
// XR: 0 = single GPU, 1 = flag protocol (post + sequence flags), 2 = LL protocol (flag-in-data push, no fences: every thread
// pushes the sums of its own elements into all ranks' buffers and collects the peers' copies of the same elements)
template <int XR>
__global__ void __launch_bounds__(256) l_finalize_kernel(int64_t m, int rc, int nparts, const double* __restrict__ part, int64_t ldlp,
                                                         int64_t part_stride, int rt, const double* __restrict__ Vi, int64_t ldvi,
                                                         const double* __restrict__ Si, int64_t ldsi, int rk, double* __restrict__ L,
                                                         int64_t ldl, P2PView v, unsigned int* ticket, int s_in_smem, LLView lv) {
    extern __shared__ double Ss[];   // [rc][rk] (when it fits the default dynamic shared memory; else S is read through the cache)
    if (Vi && s_in_smem) {
        for (int e = threadIdx.x; e < rc * rk; e += blockDim.x) Ss[e] = Si[(e / rk) + (int64_t)(e % rk) * ldsi];
    }
    __syncthreads();
    const int64_t total = m * rc;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    auto init_term = [&](int64_t j, int c) {
        double t = 0.0;
        if (Vi) {
            if (s_in_smem) {
                const double* sr = Ss + c * rk;
                for (int k = 0; k < rk; ++k) t = fma(Vi[j + (int64_t)k * ldvi], sr[k], t);
            } else {
                for (int k = 0; k < rk; ++k) t = fma(Vi[j + (int64_t)k * ldvi], __ldg(Si + c + (int64_t)k * ldsi), t);
            }
        }
        return t;
    };
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gstride) {
        const int64_t j = e % m;
        const int c = (int)(e / m);
        double s = 0.0;
        // factor column c lives in the partial of cluster rank c / rt (rt columns per CTA), local column c % rt
        const double* pp = part + (int64_t)(c / rt) * ldlp * rt + j + (int64_t)(c % rt) * ldlp;
#pragma unroll 16
        for (int p = 0; p < nparts; ++p) s += pp[(int64_t)p * part_stride];
        if (XR == 2) ll_push(lv, e, s);
        else if (XR == 1) v.data_local[e] = s;
        else L[j + (int64_t)c * ldl] = s + init_term(j, c);
    }
    if (XR == 0) return;
    if (XR == 2) {
        ll_flush(lv);
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gstride) {
            const int64_t j = e % m;
            const int c = (int)(e / m);
            L[j + (int64_t)c * ldl] = ll_sum(lv, e) + init_term(j, c);
        }
        return;
    }
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();   // one cumulative fence per CTA (the barrier ordered the other threads' stores before it)
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < v.nranks) st_release_sys(v.flags_peer[threadIdx.x] + (size_t)v.rank * P2P_FLAG_STRIDE, v.seq);
        if (threadIdx.x == 0) *ticket = 0;
    }
    p2p_wait_all(v);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gstride) {
        const int64_t j = e % m;
        const int c = (int)(e / m);
        double s = 0.0;
        for (int g = 0; g < v.nranks; ++g) s += ld_relaxed_sys(v.data_peer[g] + e);
        L[j + (int64_t)c * ldl] = s + init_term(j, c);
    }
}

// Lout chunk (m x rc) = Σ_parts (+ Σ_ranks) + Vi·Siᵀ ; picks the fused kernel when it can (single GPU or P2P transport).
// Partials come from a pass launched with `rt` factor columns per CTA and `csize` CTAs per cluster, `nparts` clusters.
inline void l_finalize(dlra_engine* e, int rc, int nparts, const double* part, int64_t ldlp, int rt, int csize, const double* Vi,
                       int64_t ldvi, const double* Si, int64_t ldsi, int rk, double* L, int64_t ldl, Ctx* on = nullptr) {
    const int64_t part_stride = ldlp * rt * csize;
    Ctx& cx = on ? *on : e->cx;   // the L-side tail may run on the auxiliary stream beside the K-side chain
    Comm& cm = e->comm;
    DLRA_REQUIRE(on == nullptr || cm.nranks <= 1 || cm.p2p, "library collectives stay on the main stream");
    const int chan = on ? 1 : 0;  // the auxiliary stream owns exchange channel 1 (channel 0: collectives of the main stream)
    const int64_t total = e->m * (int64_t)rc;
    // the cross-rank variant spins on peer flags, so its whole grid must be co-resident (256 threads, <= 16 KB smem: >= 4 CTAs/SM)
    const bool xr = cm.nranks > 1 && cm.p2p;
    // (on the auxiliary stream the spinning CTAs must leave whole SMs to the K-side TSQR that runs beside them)
    const int grid = (int)std::min<int64_t>(cdiv(total, 256), xr ? (on ? (int64_t)cx.num_sms / 2 : 4 * (int64_t)cx.num_sms) : ((int64_t)1 << 30));
    const int s_in_smem = (size_t)rc * rk * sizeof(double) <= 40 * 1024 ? 1 : 0;
    const size_t smem = (Vi && s_in_smem) ? (size_t)rc * rk * sizeof(double) : 0;
    if (cm.nranks <= 1) {
        l_finalize_kernel<0><<<grid, 256, smem, cx.stream>>>(e->m, rc, nparts, part, ldlp, part_stride, rt, Vi, ldvi, Si, ldsi, rk, L, ldl, P2PView{}, nullptr, s_in_smem, LLView{});
        cx.launches++;
    } else if (cm.ll_fits(total)) {
        l_finalize_kernel<2><<<grid, 256, smem, cx.stream>>>(e->m, rc, nparts, part, ldlp, part_stride, rt, Vi, ldvi, Si, ldsi, rk, L, ldl, P2PView{}, nullptr, s_in_smem, cm.next_ll(chan));
        cx.launches++;
    } else if (cm.p2p) {
        DLRA_REQUIRE((size_t)total * 8 <= cm.xdata_bytes, "P2P exchange region too small for an L chunk");
        P2PView v = cm.next_view(chan);
        l_finalize_kernel<1><<<grid, 256, smem, cx.stream>>>(e->m, rc, nparts, part, ldlp, part_stride, rt, Vi, ldvi, Si, ldsi, rk, L, ldl, v, cm.ticket_of(chan), s_in_smem, LLView{});
        cx.launches++;
    } else {
        // NCCL transport: local reduction, library all-reduce of the dense chunk (ldl == m), then the initial term
        for (int q = 0; q * rt < rc; ++q)
            reduce_parts(cx, (int)e->m, std::min(rt, rc - q * rt), nparts, part + (int64_t)q * ldlp * rt, ldlp, part_stride,
                         L + (int64_t)q * rt * ldl, ldl, 1.0, 0.0);
        if (ldl == e->m) cm.allreduce_sum(L, total, cx);
        else for (int c = 0; c < rc; ++c) cm.allreduce_sum(L + (int64_t)c * ldl, e->m, cx);
        if (Vi) gemm_nn(cx, e->m, rk, rc, Vi, ldvi, nullptr, 0, Si, ldsi, true, L, ldl, 1.0, 1.0);
    }
    DLRA_CUDA(cudaGetLastError());
}

// tuning knobs for A/B runs on hardware (read once): DLRA_MAX_CLUSTER = 1..4 CTAs per cluster (1 restores one sweep per 16
// factor columns), DLRA_KONLY_RT = 16 | 32 factor columns per CTA in K-only sweeps wider than 16 columns
inline int env_int(const char* name, int dflt, int lo, int hi) {
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    const int v = atoi(s);
    return v < lo ? lo : (v > hi ? hi : v);
}
// measured on B200 (profiles/wide_rank_r02.txt): pairs use all 148 SMs (74 co-resident clusters); clusters of 4 leave ~10 % of the SMs idle
inline int pass_max_cluster() { static int v = env_int("DLRA_MAX_CLUSTER", 2, 1, PT_MAX_CLUSTER); return v; }
inline int pass_konly_rt() { static int v = env_int("DLRA_KONLY_RT", 32, 16, 32); return v >= 32 ? 32 : 16; }
inline int pass_fused_rt() { static int v = env_int("DLRA_FUSED_RT", 16, 16, 32); return v >= 32 ? 32 : 16; }

// K (n x r) += ΔA·Vf and/or Lout (m x r, ldl) = ΔAᵀ·Uf.  One sweep over ΔA covers up to 32·C factor columns: a cluster of
// up to C CTAs, 16 or 32 columns each, shares every ΔA tile through TMA multicast; wider blocks take several sweeps.
// Lout is COMPLETE on return: summed over this rank's panels, over the ranks of a row-sharded run, plus Vi·Siᵀ if given.
inline void tma_pass_KL(dlra_engine* e, const Delta& d, int r, const double* Vf, int64_t ldv, const double* Uf, int64_t ldu,
                        double* K, int64_t ldk, double* Lout, int64_t ldl, const double* Vi = nullptr, int64_t ldvi = 0,
                        const double* Si = nullptr, int64_t ldsi = 0, const PassOpts& po = PassOpts()) {
    // factor operands must satisfy the TMA alignment rules too; otherwise stage them through aligned scratch
    DLRA_REQUIRE((!K || tma_ok(Vf, ldv)) && (!Lout || tma_ok(Uf, ldu)), "factor buffers must be 16-byte aligned with even ld");
    const int64_t ldlp = round_up(e->m, 2);
    const bool dk = K != nullptr, dl = Lout != nullptr, df = d.Aprev != nullptr;
    const int maxc = pass_max_cluster();
    for (int c0 = 0; c0 < r;) {
        const int rem = r - c0;
        double* Kc = K ? K + (int64_t)c0 * ldk : nullptr;
        const double* Vc = Vf ? Vf + (int64_t)c0 * ldv : nullptr;
        const double* Uc = Uf ? Uf + (int64_t)c0 * ldu : nullptr;
        const int rt = rem <= 8 ? 8 : (rem <= 16 ? 16 : ((dk && !dl) ? pass_konly_rt() : pass_fused_rt()));
        const int csize = std::min(maxc, (int)cdiv(rem, rt));
        const int rc = std::min(rem, rt * csize);
        PassGrid g;
        if (rt == 8) g = pass_grid_rt<8>(e, dk, dl, df, csize);
        else if (rt == 16) g = pass_grid_rt<16>(e, dk, dl, df, csize);
        else g = pass_grid_rt<32>(e, dk, dl, df, csize);
        if (dl) e->part.ensure((int64_t)g.ctas() * ldlp * rt, e->cx.stream);
        double* Lp = dl ? e->part.p : nullptr;
        if (rt == 8) launch_pass_rt<8>(e, d, rc, Vc, ldv, Uc, ldu, Kc, ldk, Lp, ldlp, g, po);
        else if (rt == 16) launch_pass_rt<16>(e, d, rc, Vc, ldv, Uc, ldu, Kc, ldk, Lp, ldlp, g, po);
        else launch_pass_rt<32>(e, d, rc, Vc, ldv, Uc, ldu, Kc, ldk, Lp, ldlp, g, po);
        // CTA q of every cluster holds the partial of the factor columns [rt*q, rt*q + rt)
        if (dl) l_finalize(e, rc, g.nclusters, Lp, ldlp, rt, csize, Vi, ldvi, Si ? Si + c0 : nullptr, ldsi, r, Lout + (int64_t)c0 * ldl, ldl);
        c0 += rc;
    }
}

// Tall product on the streaming machinery:  C (n x q) = beta*C + alpha * A (n x p) * op(B),  beta in {0, 1} — the tall factor A
// plays the role of ΔA in a K-only sweep (TMA tiles, DMMA), op(B) is staged (scaled, transposed) into a dense p x q block.
// Used for the O(n r^2) products of wide steps (U0*S0, the BCGS2 projections, Uhat*P), where the one-thread-per-row kernel
// runs at a third of this speed.  Returns false when the shape or alignment does not fit (caller falls back).
inline bool tall_gemm_tma(dlra_engine* e, int64_t n, int p, int q, const double* A, int64_t lda, const double* B, int64_t ldb, bool transB,
                          double* C, int64_t ldc, double alpha, double beta) {
    if (e->flags & DLRA_FORCE_GENERIC) return false;
    // measured (profiles/r02): pays off for wide products only — thin ones (BCGS2 projections with q = 16, ranks <= 32) are
    // faster on the one-thread-per-row kernel, whose launch has no pipeline to fill
    if (n != e->n || n < 65536 || p < 48 || q < 32 || (p % 2) != 0 || (beta != 0.0 && beta != 1.0)) return false;
    if (n >= ((int64_t)1 << 31) || (n % 2) != 0 || !tma_ok(A, lda)) return false;
    Ctx& cx = e->cx;
    e->bstage.ensure((int64_t)p * q, cx.stream);
    copy_mat(cx, p, q, B, ldb, transB, e->bstage.p, p, alpha, 0.0);
    Delta d;
    d.A = A; d.lda = lda;
    PassOpts po;
    po.mcols = p; po.kstore = (beta == 0.0); po.timed = false;
    tma_pass_KL(e, d, q, e->bstage.p, p, nullptr, 0, C, ldc, nullptr, 0, nullptr, 0, nullptr, 0, po);
    return true;
}

// Sout (p x q) = Lfᵀ·(ΔA·Rf): the K-use streams W = ΔA·Rf (n x q, 8·n·q bytes — <1% of the ΔA traffic) and a
// tall-skinny Gram product folds it with Lf.
inline void tma_pass_S(dlra_engine* e, const Delta& d, int p, int q, const double* Lf, int64_t ldlf, const double* Rf, int64_t ldrf,
                       double* Sout, int64_t lds) {
    Ctx& cx = e->cx;
    e->nscr.ensure(e->n * (int64_t)q, cx.stream);
    DLRA_CUDA(cudaMemsetAsync(e->nscr.p, 0, (size_t)e->n * q * 8, cx.stream));
    tma_pass_KL(e, d, q, Rf, ldrf, nullptr, 0, e->nscr.p, e->n, nullptr, 0);
    e->gws.ensure(gemm_tn_ws(cx, e->n, p, q), cx.stream);
    gemm_tn(cx, e->n, p, q, Lf, ldlf, nullptr, 0, e->nscr.p, e->n, Sout, lds, 1.0, 0.0, e->gws.p);
}

}  // namespace dlra
