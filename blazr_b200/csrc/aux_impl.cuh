// aux_impl.cuh -- per-format templates of the one-off / test-hook kernels (repack, dense dequantize, integer partials).
#pragma once
#include "formats.cuh"
#include "internal.h"

namespace b200q {

// ------------------------------------------------------------------------------------------------
// ggml raw blocks -> tiles.  One CTA (128 threads) per chunk, thread == row.
// ------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) repack_ggml_kernel(const uint8_t* __restrict__ src, int64_t src_row_bytes, int64_t n0, int64_t k0,
                                                           int64_t N, int64_t K, int64_t KC, uint8_t* __restrict__ dst, FmtMeta meta) {
    int64_t kc = blockIdx.x, t = blockIdx.y;
    int r = threadIdx.x;
    uint8_t* chunk = dst + (t * KC + kc) * (int64_t)F::chunk_bytes(meta.gpc);
    int64_t nl = t * TILE_ROWS + r;
    constexpr int BE = F::src_block_elems(), BB = F::src_block_bytes();
    int64_t kbeg = kc * CHUNK_K;
    int nvalid = 0;
    if (nl < N && kbeg < K) {
        int64_t rem = (K - kbeg) / BE;
        int per_chunk = CHUNK_K / BE;
        nvalid = (int)(rem < per_chunk ? rem : per_chunk);
    }
    const uint8_t* s = src + (n0 + (nl < N ? nl : 0)) * src_row_bytes + ((k0 + kbeg) / BE) * BB;
    F::repack_row(s, nvalid, chunk, r, meta);
}

// ------------------------------------------------------------------------------------------------
// dense dequantize: out[N,K] = a * (v - off) - b  (separate multiply / subtract: bit-exact contract #1)
// ------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256) dequant_kernel(const uint8_t* __restrict__ data, int64_t N, int64_t K, int64_t KC, void* out, int dtype,
                                                       const int32_t* __restrict__ perm, FmtMeta meta) {
    int64_t kc = blockIdx.x, t = blockIdx.y;
    const uint8_t* chunk = data + (t * KC + kc) * (int64_t)F::chunk_bytes(meta.gpc);
    for (int item = threadIdx.x; item < TILE_ROWS * 8; item += blockDim.x) {
        int r = item >> 3, i = item & 7;
        int64_t n = t * TILE_ROWS + r;
        if (n >= N) continue;
        Unit u;
        F::template load_unit<false>(chunk, r, i, u, meta);
#pragma unroll
        for (int e = 0; e < 32; e++) {
            int64_t kp = kc * CHUNK_K + 32 * i + e;
            if (kp >= K) continue;
            int h = e >> 4;
            int q = unit_elem(u, e) - u.off[h];
            float v = __fsub_rn(__fmul_rn(u.a[h], (float)q), u.b[h]);
            int64_t k = perm ? perm[kp] : kp;
            store_out(out, dtype, n * K + k, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// integer partials dump (bit-exact contract #2): out[m][n][p] = sum_{k in sub-block p} (v - off) * xq
// ------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256) int_partials_kernel(const uint8_t* __restrict__ data, const uint8_t* __restrict__ xq, int64_t N,
                                                            int64_t M, int64_t KC, int32_t* __restrict__ out, FmtMeta meta) {
    int64_t kc = blockIdx.x, t = blockIdx.y, m = blockIdx.z;
    const uint8_t* chunk = data + (t * KC + kc) * (int64_t)F::chunk_bytes(meta.gpc);
    const uint8_t* rec = xq + (kc * M + m) * ACT_REC_BYTES;
    int64_t P = KC * (CHUNK_K / F::SUB);
    for (int item = threadIdx.x; item < TILE_ROWS * 8; item += blockDim.x) {
        int r = item >> 3, i = item & 7;
        int64_t n = t * TILE_ROWS + r;
        if (n >= N) continue;
        Unit u;
        F::template load_unit<false>(chunk, r, i, u, meta);
        const int32_t* xw = reinterpret_cast<const int32_t*>(rec + 32 * i);
        int sA = 0, sB = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { sA = __dp4a((int)u.v[k], xw[k], sA); sB = __dp4a((int)u.v[4 + k], xw[4 + k], sB); }
        uint32_t bs = reinterpret_cast<const uint32_t*>(rec + 288)[i];
        int bA = (int)(int16_t)(bs & 0xFFFF), bB = (int)(int16_t)(bs >> 16);
        sA -= u.off[0] * bA;
        sB -= u.off[1] * bB;
        int32_t* o = out + (m * N + n) * P;
        if (F::SUB == 32) o[kc * 8 + i] = sA + sB;
        else { o[kc * 16 + 2 * i] = sA; o[kc * 16 + 2 * i + 1] = sB; }
    }
}


template <int FAMILY>
cudaError_t repack_launch(const uint8_t* src, int64_t src_row_bytes, int64_t n0, int64_t k0, const b200q_weight* w, cudaStream_t st);
template <int FAMILY>
cudaError_t dequant_launch(const b200q_weight* w, void* out, int dtype, cudaStream_t st);
template <int FAMILY>
cudaError_t partials_launch(const b200q_weight* w, const uint8_t* xq, int64_t M, int32_t* out, cudaStream_t st);

}  // namespace b200q
