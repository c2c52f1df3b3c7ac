// inst_body.cuh -- instantiates every per-format kernel for one format (included by inst_<format>.cu with
// B200Q_FMT / B200Q_FAM_ID / B200Q_HAS_GGML_REPACK defined), so formats compile as parallel translation units.
#include "aux_impl.cuh"
#include "gemm_impl.cuh"
#include "matvec_impl.cuh"

namespace b200q {

template <>
cudaError_t mv_launch<B200Q_FAM_ID>(const MatvecParams& p, int mb, int grid, int smem, cudaStream_t st) {
    return launch_f<B200Q_FMT>(p, mb, grid, smem, st);
}
template <>
cudaError_t gemm_launch<B200Q_FAM_ID>(const GemmParams& p, int grid, int smem, cudaStream_t st) {
    return launch_gemm_t<B200Q_FMT>(p, grid, smem, st);
}
template <>
cudaError_t dequant_launch<B200Q_FAM_ID>(const b200q_weight* w, void* out, int dtype, cudaStream_t st) {
    dim3 grid((unsigned)w->KC, (unsigned)w->T);
    dequant_kernel<B200Q_FMT><<<grid, 256, 0, st>>>(w->data, w->N, w->K, w->KC, out, dtype, w->perm, FmtMeta{w->gpc});
    count_launch();
    return cudaGetLastError();
}
template <>
cudaError_t partials_launch<B200Q_FAM_ID>(const b200q_weight* w, const uint8_t* xq, int64_t M, int32_t* out, cudaStream_t st) {
    dim3 grid((unsigned)w->KC, (unsigned)w->T, (unsigned)M);
    int_partials_kernel<B200Q_FMT><<<grid, 256, 0, st>>>(w->data, xq, w->N, M, w->KC, out, FmtMeta{w->gpc});
    count_launch();
    return cudaGetLastError();
}
#if B200Q_HAS_GGML_REPACK
template <>
cudaError_t repack_launch<B200Q_FAM_ID>(const uint8_t* src, int64_t src_row_bytes, int64_t n0, int64_t k0, const b200q_weight* w, cudaStream_t st) {
    dim3 grid((unsigned)w->KC, (unsigned)w->T);
    repack_ggml_kernel<B200Q_FMT><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, FmtMeta{w->gpc});
    count_launch();
    return cudaGetLastError();
}
#endif

}  // namespace b200q
