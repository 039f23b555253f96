// placeholder until the TMA + DMMA streaming kernels land
#pragma once
#include "engine.cuh"
namespace dlra {
inline bool tma_pass_supported(int64_t, int64_t, const Delta&) { return false; }
inline void tma_pass_KL(dlra_engine*, const Delta&, int, const double*, int64_t, const double*, int64_t, double*, int64_t, double*, int64_t) {}
inline void tma_pass_S(dlra_engine*, const Delta&, int, int, const double*, int64_t, const double*, int64_t, double*, int64_t) {}
}  // namespace dlra
