// placeholder: DE-problem sub-flows land next
#pragma once
#include "engine.cuh"
namespace dlra {
inline void de_release(dlra_engine*) {}
inline void de_rank_changed(dlra_engine*) {}
inline void de_rhs_set(dlra_engine*, const dlra_operator*, const dlra_operator*, const double*, int64_t, const double*, int64_t, int,
                       const dlra_operator*, const dlra_operator*, double) { throw CudaError(6, "DE right-hand sides not built yet"); }
inline void de_K_flow(dlra_engine*, double*, int, const double*, double, double) { throw CudaError(6, "DE flows not built yet"); }
inline void de_L_flow(dlra_engine*, double*, int, const double*, double, double) { throw CudaError(6, "DE flows not built yet"); }
inline void de_S_flow(dlra_engine*, double*, int, int, const double*, const double*, double, double, double) { throw CudaError(6, "DE flows not built yet"); }
}  // namespace dlra
