// NCCL plumbing for the row-sharded multi-GPU path.  The reference is single-process and has
// no collective; here L / S / M partial sums are all-reduced and the TSQR R-factors all-gathered.
// libnccl.so.2 is dlopen'ed at dlra_comm_init time (the torch-bundled copy is reused when the host process
// already loaded it), so libdlra.so itself has no link-time NCCL dependency and loads on CPU-only boxes.
#pragma once
#include "common.cuh"
#include <dlfcn.h>

namespace dlra {

struct Comm {
    struct UniqueId { char internal[128]; };
    int nranks = 1, rank = 0;
    void* lib = nullptr;
    void* comm = nullptr;  // ncclComm_t
    // ncclResult_t (*)(...)
    int (*pGetUniqueId)(void*) = nullptr;
    int (*pCommInitRank)(void**, int, /*ncclUniqueId by value*/ UniqueId, int) = nullptr;
    int (*pCommDestroy)(void*) = nullptr;
    int (*pAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*pAllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*pGetErrorString)(int) = nullptr;

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
    }
    void destroy() {
        if (comm && pCommDestroy) pCommDestroy(comm);
        comm = nullptr;
    }
    // in-place sum over ranks (ncclDouble = 8, ncclSum = 0)
    void allreduce_sum(double* buf, int64_t count, cudaStream_t s) {
        if (nranks <= 1 || count <= 0) return;
        check(pAllReduce(buf, buf, (size_t)count, 8, 0, comm, s), "ncclAllReduce");
    }
    void allgather(const double* send, double* recv, int64_t count_per_rank, cudaStream_t s) {
        if (nranks <= 1) {
            if (send != recv) DLRA_CUDA(cudaMemcpyAsync(recv, send, count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, s));
            return;
        }
        check(pAllGather(send, recv, (size_t)count_per_rank, 8, comm, s), "ncclAllGather");
    }
};

}  // namespace dlra
