// tcgen05 (5th-gen tensor core) implicit-GEMM kernels.  Entry points return false when they do
// not handle a case; the engine then runs the CUDA-core kernel of igemm_simt.cuh instead.
#pragma once
#include "common.cuh"

namespace ddpm {
namespace tc {

inline bool available() { return false; }
inline void init() {}

template <typename TIn, typename TOut>
bool conv3x3(cudaStream_t, const TIn*, int, const TIn*, int, const TIn*, int, TOut*, const Geo&, const float*,
             const float*, int, double*) {
    return false;
}
template <typename TA>
bool up2(cudaStream_t, const TA*, const TA*, TA*, const Geo&, const Geo&, const float*) {
    return false;
}
template <typename TG, typename TA>
bool wgrad3x3(cudaStream_t, const TG*, int, const TA*, int, const Geo&, float*, int, int) {
    return false;
}

}  // namespace tc
}  // namespace ddpm
