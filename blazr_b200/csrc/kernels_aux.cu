// kernels_aux.cu -- one-off and test-hook kernels: upload repack, dense dequantize, activation quantiser,
// integer-partial dump, activation -> bf16 staging.
#include "aux_impl.cuh"

namespace b200q {

cudaError_t launch_repack_ggml(int family, const uint8_t* src, int64_t src_row_bytes, int64_t n0, int64_t k0, const b200q_weight* w,
                               cudaStream_t st) {
    (void)family;
    // source adaptors (formats.cuh): ggml types re-encoded into an existing family
    if (w->ggml_type == 2 || w->ggml_type == 6 || w->ggml_type == 20 || w->ggml_type == 34 || (w->ggml_type >= 16 && w->ggml_type <= 18) || w->ggml_type == 21 || w->ggml_type == 22 || w->ggml_type == 19 || w->ggml_type == 29) {
        dim3 grid((unsigned)w->KC, (unsigned)w->T);
        const FmtMeta meta{w->gpc};
        if (w->ggml_type == 2) repack_ggml_kernel<SrcQ4_0><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        else if (w->ggml_type == 6) repack_ggml_kernel<SrcQ5_0><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        else if (w->ggml_type == 16) repack_ggml_kernel<SrcIQ2XXS><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        else if (w->ggml_type == 17) repack_ggml_kernel<SrcIQ2XS><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        else if (w->ggml_type == 18) repack_ggml_kernel<SrcIQ3XXS><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        else if (w->ggml_type == 22) repack_ggml_kernel<SrcIQ2S><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        else if (w->ggml_type == 21) repack_ggml_kernel<SrcIQ3S><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        else if (w->ggml_type == 19) repack_ggml_kernel<SrcIQ1S><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        else if (w->ggml_type == 29) repack_ggml_kernel<SrcIQ1M><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        else if (w->ggml_type == 34) repack_ggml_kernel<SrcTQ1_0><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        else repack_ggml_kernel<SrcIQ4NL><<<grid, 128, 0, st>>>(src, src_row_bytes, n0, k0, w->N, w->K, w->KC, w->data, meta);
        count_launch();
        return cudaGetLastError();
    }
    switch (w->family) {
        case B200Q_FAM_Q4_K: return repack_launch<B200Q_FAM_Q4_K>(src, src_row_bytes, n0, k0, w, st);
        case B200Q_FAM_Q6_K: return repack_launch<B200Q_FAM_Q6_K>(src, src_row_bytes, n0, k0, w, st);
        case B200Q_FAM_Q8_0: return repack_launch<B200Q_FAM_Q8_0>(src, src_row_bytes, n0, k0, w, st);
        case B200Q_FAM_Q5_K: return repack_launch<B200Q_FAM_Q5_K>(src, src_row_bytes, n0, k0, w, st);
        case B200Q_FAM_Q4_1: return repack_launch<B200Q_FAM_Q4_1>(src, src_row_bytes, n0, k0, w, st);
        case B200Q_FAM_Q5_1: return repack_launch<B200Q_FAM_Q5_1>(src, src_row_bytes, n0, k0, w, st);
        case B200Q_FAM_Q2_K: return repack_launch<B200Q_FAM_Q2_K>(src, src_row_bytes, n0, k0, w, st);
        case B200Q_FAM_Q3_K: return repack_launch<B200Q_FAM_Q3_K>(src, src_row_bytes, n0, k0, w, st);
        case B200Q_FAM_IQ4_XS: return repack_launch<B200Q_FAM_IQ4_XS>(src, src_row_bytes, n0, k0, w, st);
        case B200Q_FAM_TQ2_0: return repack_launch<B200Q_FAM_TQ2_0>(src, src_row_bytes, n0, k0, w, st);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// AWQ (reference src/loader/safetensors/awq.rs:29-32,190-226): qweight u32 [K, N/8] with nibble order
// AWQ_SHIFTS, scales f32 [K/gs, N], zeros f32 [K/gs, N] (already unpacked).
// ------------------------------------------------------------------------------------------------
__device__ __constant__ uint32_t kAwqShifts[8] = {0, 16, 4, 20, 8, 24, 12, 28};

__global__ void __launch_bounds__(128) repack_awq_kernel(const uint32_t* __restrict__ qweight, const float* __restrict__ scales,
                                                          const float* __restrict__ zeros, int64_t N_full, int64_t n0, int64_t k0, int64_t N,
                                                          int64_t K, int64_t KC, int gs, int gpc, uint8_t* __restrict__ dst, int* err) {
    int64_t kc = blockIdx.x, t = blockIdx.y;
    int r = threadIdx.x;
    uint8_t* chunk = dst + (t * KC + kc) * (int64_t)FmtG4::chunk_bytes(gpc);
    int64_t nl = t * TILE_ROWS + r;
    uint8_t q[256];
    uint16_t sc[8];
    uint8_t z[8];
    for (int g = 0; g < gpc; g++) { sc[g] = 0; z[g] = 0; }
    bool row_ok = nl < N;
    int64_t n = n0 + (row_ok ? nl : 0);
    int64_t n8 = N_full / 8;
    uint32_t sh = kAwqShifts[n & 7];
    for (int j = 0; j < 256; j++) {
        int64_t kl = kc * CHUNK_K + j;
        uint32_t v = 0;
        if (row_ok && kl < K) v = (qweight[(k0 + kl) * n8 + (n >> 3)] >> sh) & 0xF;
        q[j] = (uint8_t)v;
    }
    int span = CHUNK_K / gpc;  // k covered by one stored group slot
    for (int g = 0; g < gpc; g++) {
        int64_t kl = kc * CHUNK_K + (int64_t)g * span;
        if (row_ok && kl < K) {
            int64_t grp = (k0 + kl) / gs;
            float s = scales[grp * N_full + n];
            float zf = zeros[grp * N_full + n];
            __half h = __float2half_rn(s);
            if (__half2float(h) != s) atomicOr(err, 1);
            int zi = (int)zf;
            if ((float)zi != zf || zi < 0 || zi > 255) atomicOr(err, 2);
            sc[g] = __half_as_ushort(h);
            z[g] = (uint8_t)zi;
        }
    }
    FmtG4::store_row(q, sc, z, chunk, r, gpc);
}

cudaError_t launch_repack_awq(const uint32_t* qweight, const float* scales, const float* zeros, int64_t N_full, int64_t n0, int64_t k0,
                              const b200q_weight* w, int* err_flag, cudaStream_t st) {
    dim3 grid((unsigned)w->KC, (unsigned)w->T);
    repack_awq_kernel<<<grid, 128, 0, st>>>(qweight, scales, zeros, N_full, n0, k0, w->N, w->K, w->KC, w->group_size, w->gpc, w->data, err_flag);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// GPTQ (reference src/loader/safetensors/gptq.rs:198-259): qweight u32 [K/8, N] sequential nibbles
// along K, scales f32 [G, N], qzeros u32 [G, N/8] sequential nibbles along N, perm = group-sorting
// permutation of K (act-order) or null.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) repack_gptq_kernel(const uint32_t* __restrict__ qweight, const float* __restrict__ scales,
                                                           const uint32_t* __restrict__ qzeros, const int32_t* __restrict__ perm, int zpo,
                                                           int64_t N_full, int64_t n0, int64_t k0, int64_t N, int64_t K, int64_t KC, int gs,
                                                           int gpc, uint8_t* __restrict__ dst, int* err) {
    int64_t kc = blockIdx.x, t = blockIdx.y;
    int r = threadIdx.x;
    uint8_t* chunk = dst + (t * KC + kc) * (int64_t)FmtG4::chunk_bytes(gpc);
    int64_t nl = t * TILE_ROWS + r;
    uint8_t q[256];
    uint16_t sc[8];
    uint8_t z[8];
    for (int g = 0; g < gpc; g++) { sc[g] = 0; z[g] = 0; }
    bool row_ok = nl < N;
    int64_t n = n0 + (row_ok ? nl : 0);
    for (int j = 0; j < 256; j++) {
        int64_t kl = kc * CHUNK_K + j;
        uint32_t v = 0;
        if (row_ok && kl < K) {
            int64_t kp = k0 + kl;
            int64_t k = perm ? perm[kp] : kp;
            v = (qweight[(k >> 3) * N_full + n] >> (4 * (k & 7))) & 0xF;
        }
        q[j] = (uint8_t)v;
    }
    int span = CHUNK_K / gpc;
    for (int g = 0; g < gpc; g++) {
        int64_t kl = kc * CHUNK_K + (int64_t)g * span;
        if (row_ok && kl < K) {
            int64_t grp = (k0 + kl) / gs;
            float s = scales[grp * N_full + n];
            __half h = __float2half_rn(s);
            if (__half2float(h) != s) atomicOr(err, 1);
            uint32_t zz = (qzeros[grp * (N_full / 8) + (n >> 3)] >> (4 * (n & 7))) & 0xF;
            sc[g] = __half_as_ushort(h);
            z[g] = (uint8_t)(zz + zpo);
        }
    }
    FmtG4::store_row(q, sc, z, chunk, r, gpc);
}

cudaError_t launch_repack_gptq(const uint32_t* qweight, const float* scales, const uint32_t* qzeros, const int32_t* perm, int zpo,
                               int64_t N_full, int64_t n0, int64_t k0, const b200q_weight* w, int* err_flag, cudaStream_t st) {
    dim3 grid((unsigned)w->KC, (unsigned)w->T);
    repack_gptq_kernel<<<grid, 128, 0, st>>>(qweight, scales, qzeros, perm, zpo, N_full, n0, k0, w->N, w->K, w->KC, w->group_size, w->gpc,
                                             w->data, err_flag);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_dequantize(const b200q_weight* w, void* out, int dtype, cudaStream_t st) {
    switch (w->family) {
        case B200Q_FAM_Q4_K: return dequant_launch<B200Q_FAM_Q4_K>(w, out, dtype, st);
        case B200Q_FAM_Q6_K: return dequant_launch<B200Q_FAM_Q6_K>(w, out, dtype, st);
        case B200Q_FAM_Q8_0: return dequant_launch<B200Q_FAM_Q8_0>(w, out, dtype, st);
        case B200Q_FAM_Q5_K: return dequant_launch<B200Q_FAM_Q5_K>(w, out, dtype, st);
        case B200Q_FAM_Q4_1: return dequant_launch<B200Q_FAM_Q4_1>(w, out, dtype, st);
        case B200Q_FAM_Q5_1: return dequant_launch<B200Q_FAM_Q5_1>(w, out, dtype, st);
        case B200Q_FAM_Q2_K: return dequant_launch<B200Q_FAM_Q2_K>(w, out, dtype, st);
        case B200Q_FAM_Q3_K: return dequant_launch<B200Q_FAM_Q3_K>(w, out, dtype, st);
        case B200Q_FAM_IQ4_XS: return dequant_launch<B200Q_FAM_IQ4_XS>(w, out, dtype, st);
        case B200Q_FAM_TQ2_0: return dequant_launch<B200Q_FAM_TQ2_0>(w, out, dtype, st);
        case B200Q_FAM_I8S: return dequant_launch<B200Q_FAM_I8S>(w, out, dtype, st);
        case B200Q_FAM_G4: return dequant_launch<B200Q_FAM_G4>(w, out, dtype, st);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// activation quantiser: per 32-element block d = amax/127, id = 1/d, q = roundf(x*id) (oracle orc_quantize_act).
// Output records, chunk-major so the matvec brings one k-chunk of all M rows with one bulk copy:
//   xq[(kc*M + m)*320] = { int8 q[256]; float d[8]; (int16 bsum16[2])[8] }
// One CTA of 256 threads per (kc, m); warp == 32-block.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) act_quant_kernel(const void* __restrict__ x, int x_dtype, int64_t M, int64_t K, int64_t ldx,
                                                         const int32_t* __restrict__ perm, uint8_t* __restrict__ xq) {
    int64_t kc = blockIdx.x, m = blockIdx.y;
    int t = threadIdx.x, wid = t >> 5, lane = t & 31;
    int64_t k = kc * CHUNK_K + t;
    pdl_launch_dependents();  // the matvec that consumes xq may start prefetching its weights now
    pdl_wait();               // x is produced by the preceding kernel; xq may still be read by an earlier matvec
    float v = 0.0f;
    if (k < K) v = load_in(x, x_dtype, m * ldx + (perm ? perm[k] : k));
    float amax = fabsf(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    float d = __fdiv_rn(amax, 127.0f);
    const float id = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
    int q = (int)roundf(__fmul_rn(v, id));
    int s = q;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);  // sum within each 16-lane half
    int s_hi = __shfl_sync(0xffffffffu, s, 16);
    uint8_t* rec = xq + (kc * M + m) * ACT_REC_BYTES;
    rec[t] = (uint8_t)(int8_t)q;
    if (lane == 0) {
        reinterpret_cast<float*>(rec + 256)[wid] = d;
        reinterpret_cast<uint32_t*>(rec + 288)[wid] = ((uint32_t)s & 0xFFFFu) | ((uint32_t)s_hi << 16);
    }
}

cudaError_t launch_act_quant(const void* x, int x_dtype, int64_t M, int64_t K, int64_t K_pad, int64_t ldx, const int32_t* perm, uint8_t* xq,
                             cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(K_pad / CHUNK_K), (unsigned)M);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, act_quant_kernel, x, x_dtype, M, K, ldx, perm, xq);
    count_launch();
    return e;
}

__global__ void act_unpack_kernel(const uint8_t* __restrict__ xq, int64_t M, int64_t K_pad, int8_t* q, float* d, int32_t* bsum16) {
    int64_t kc = blockIdx.x, m = blockIdx.y;
    int t = threadIdx.x;
    const uint8_t* rec = xq + (kc * M + m) * ACT_REC_BYTES;
    q[m * K_pad + kc * CHUNK_K + t] = (int8_t)rec[t];
    if (t < 8) {
        d[m * (K_pad / 32) + kc * 8 + t] = reinterpret_cast<const float*>(rec + 256)[t];
        uint32_t bs = reinterpret_cast<const uint32_t*>(rec + 288)[t];
        bsum16[m * (K_pad / 16) + kc * 16 + 2 * t] = (int)(int16_t)(bs & 0xFFFF);
        bsum16[m * (K_pad / 16) + kc * 16 + 2 * t + 1] = (int)(int16_t)(bs >> 16);
    }
}
cudaError_t launch_act_unpack(const uint8_t* xq, int64_t M, int64_t K_pad, int8_t* q, float* d, int32_t* bsum16, cudaStream_t st) {
    dim3 grid((unsigned)(K_pad / CHUNK_K), (unsigned)M);
    act_unpack_kernel<<<grid, 256, 0, st>>>(xq, M, K_pad, q, d, bsum16);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_int_partials(const b200q_weight* w, const uint8_t* xq, int64_t M, int32_t* out, cudaStream_t st) {
    switch (w->family) {
        case B200Q_FAM_Q4_K: return partials_launch<B200Q_FAM_Q4_K>(w, xq, M, out, st);
        case B200Q_FAM_Q6_K: return partials_launch<B200Q_FAM_Q6_K>(w, xq, M, out, st);
        case B200Q_FAM_Q8_0: return partials_launch<B200Q_FAM_Q8_0>(w, xq, M, out, st);
        case B200Q_FAM_Q5_K: return partials_launch<B200Q_FAM_Q5_K>(w, xq, M, out, st);
        case B200Q_FAM_Q4_1: return partials_launch<B200Q_FAM_Q4_1>(w, xq, M, out, st);
        case B200Q_FAM_Q5_1: return partials_launch<B200Q_FAM_Q5_1>(w, xq, M, out, st);
        case B200Q_FAM_Q2_K: return partials_launch<B200Q_FAM_Q2_K>(w, xq, M, out, st);
        case B200Q_FAM_Q3_K: return partials_launch<B200Q_FAM_Q3_K>(w, xq, M, out, st);
        case B200Q_FAM_IQ4_XS: return partials_launch<B200Q_FAM_IQ4_XS>(w, xq, M, out, st);
        case B200Q_FAM_TQ2_0: return partials_launch<B200Q_FAM_TQ2_0>(w, xq, M, out, st);
        case B200Q_FAM_I8S: return partials_launch<B200Q_FAM_I8S>(w, xq, M, out, st);
        case B200Q_FAM_G4: return partials_launch<B200Q_FAM_G4>(w, xq, M, out, st);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// activations -> bf16 [M_pad, K_pad] (zero padded, optional K permutation) for the tcgen05 GEMM
// ------------------------------------------------------------------------------------------------
__global__ void to_bf16_kernel(const void* __restrict__ x, int x_dtype, int64_t M, int64_t K, int64_t K_pad, int64_t ldx,
                               const int32_t* __restrict__ perm, __nv_bfloat16* __restrict__ out) {
    int64_t m = blockIdx.y;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < K_pad; k += (int64_t)gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (m < M && k < K) v = load_in(x, x_dtype, m * ldx + (perm ? perm[k] : k));
        out[m * K_pad + k] = __float2bfloat16_rn(v);
    }
}
cudaError_t launch_to_bf16(const void* x, int x_dtype, int64_t M, int64_t K, int64_t K_pad, int64_t M_pad, int64_t ldx, const int32_t* perm,
                           void* out_bf16, cudaStream_t st) {
    unsigned gx = (unsigned)((K_pad + 255) / 256);
    if (gx > 64) gx = 64;
    dim3 grid(gx, (unsigned)M_pad);
    to_bf16_kernel<<<grid, 256, 0, st>>>(x, x_dtype, M, K, K_pad, ldx, perm, reinterpret_cast<__nv_bfloat16*>(out_bf16));
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200q
