// Generic layer-wise kernels for descriptors the fused family does not cover
// (arbitrary widths / depth / activation).  Placeholder: filled in after the fused path.
#pragma once
#include <cuda_runtime.h>
#include "../../include/eigenpde_b200.h"

namespace epde {
inline int generic_workspace_bytes(const epde_desc*, int64_t, size_t*) { return EPDE_ERR_UNSUPPORTED; }
inline int generic_partial(const epde_desc*, const float*, const float*, const int64_t*, int64_t, const float*,
                           const float*, void*, double*, cudaStream_t) { return EPDE_ERR_UNSUPPORTED; }
inline int generic_finish(const epde_desc*, const float*, const float*, const int64_t*, int64_t, const float*,
                          const float*, double, double, const double*, const double*, void*, double*, float*,
                          cudaStream_t) { return EPDE_ERR_UNSUPPORTED; }
inline int generic_forward(const epde_desc*, const float*, const int64_t*, int64_t, const float*, void*, float*,
                           cudaStream_t) { return EPDE_ERR_UNSUPPORTED; }
}  // namespace epde
