// dual-format matvec: Q4_K first weight + Q6_K second weight in one stream-K grid (q|k + v of Q4_K_M files; MatvecParams::w2)
#include "matvec_impl.cuh"

namespace b200q {
cudaError_t mv_launch_dual_q4k_q6k(const MatvecParams& p, int mb, int grid, int smem, cudaStream_t st) {
    return launch_dual<FmtQ4K, FmtQ6K>(p, mb, grid, smem, st);
}
}  // namespace b200q
