// gemm_tc.cu -- placeholder until the tcgen05 dequant-GEMM lands (returns "not supported": no fallback).
#include "internal.h"
namespace b200q {
size_t gemm_ws_bytes(const b200q_weight*, int64_t) { return 0; }
cudaError_t launch_gemm_tc(const b200q_weight*, const void*, int, int64_t, int64_t, void*, int, int64_t, uint8_t*, size_t, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace b200q
