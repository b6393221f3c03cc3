// tcgen05 / TMA GEMM engine (placeholder until the tensor-core path lands).
#pragma once
#include "gemm_generic.cuh"
namespace dic {
inline bool tc_gemm_eligible(const GemmArgs&) { return false; }
inline int tc_gemm(const GemmArgs&, cudaStream_t) { DIC_FAIL(-5, "tcgen05 engine not built"); }
}  // namespace dic
