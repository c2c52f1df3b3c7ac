// matvec.cu -- decode-time dequant-matvec (M = 1..4): the HBM-bound hot kernel.
//
// Design (DESIGN.md "Kernel 1"):
//   * stream-K over the tile-major chunk array: CTA g owns the contiguous chunk range
//     [g*C/G, (g+1)*C/G)  -> every SM streams one contiguous region of HBM, perfect balance for any N,K;
//   * one producer thread feeds a ring of shared-memory stages with cp.async.bulk (TMA engine, 1-D,
//     mbarrier complete_tx): 1 copy of the 18-35 KB weight chunk + 1 copy of the M x 320 B quantised
//     activation record per stage, so ~100 KB per SM are in flight independent of occupancy;
//   * 8 consumer warps: warp w owns rows 16w..16w+15 of the tile, 8 lanes x 16 B walk one row's
//     256-k chunk (unit i = 32 consecutive k), 4 rows per step; activations stay in registers for the
//     16 rows; integer dot products with dp4a, per-sub-block scales applied in f32;
//   * at tile end an 8-lane shuffle reduction; tiles split between CTAs are combined DETERMINISTICALLY
//     through per-CTA partial slots in the workspace: the last CTA to arrive (atomic counter) sums the
//     partials in CTA order and writes y (no float atomics, no inter-CTA waiting).
#include "formats.cuh"
#include "internal.h"

namespace b200q {

constexpr int MV_CONSUMER_WARPS = 8;
constexpr int MV_THREADS = (MV_CONSUMER_WARPS + 1) * 32;
constexpr int MV_MAX_STAGES = 8;
constexpr int MV_HDR_BYTES = 256;  // barriers + flags

struct MatvecParams {
    const uint8_t* w;
    const uint8_t* xq;
    void* y;
    const float* bias;
    float* ws_part;
    unsigned int* ws_cnt;
    int64_t N;
    int M, y_dtype;
    int64_t ldy;
    int64_t KC, C;
    int gpc, nstages, chunk_bytes, stage_bytes;
};

__device__ __forceinline__ int64_t sk_begin(int64_t g, int64_t C, int64_t G) { return g * C / G; }
__device__ __forceinline__ int64_t sk_owner(int64_t c, int64_t C, int64_t G) { return ((c + 1) * G - 1) / C; }

template <class F, int MB>
__global__ void __launch_bounds__(MV_THREADS, (MB <= 2 ? 2 : 1)) matvec_kernel(const MatvecParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + MV_MAX_STAGES;
    volatile int* sflag = reinterpret_cast<volatile int*>(smem + 2 * MV_MAX_STAGES * 8);
    uint8_t* stages = smem + MV_HDR_BYTES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t G = gridDim.x, g = blockIdx.x;
    const int64_t c0 = sk_begin(g, p.C, G), c1 = sk_begin(g + 1, p.C, G);
    const int nst = p.nstages;

    if (tid == 0) {
        for (int s = 0; s < nst; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MV_CONSUMER_WARPS);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (warp == MV_CONSUMER_WARPS) {
        // ===================== producer =====================
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            const uint32_t xbytes = (uint32_t)p.M * ACT_REC_BYTES;
            int64_t it = 0;
            for (int64_t c = c0; c < c1; c++, it++) {
                int s = (int)(it % nst);
                uint32_t round = (uint32_t)(it / nst);
                mbar_wait(&empty[s], (round & 1u) ^ 1u);
                uint8_t* st = stages + (size_t)s * p.stage_bytes;
                mbar_arrive_expect_tx(&full[s], (uint32_t)p.chunk_bytes + xbytes);
                bulk_g2s_hint(st, p.w + c * (int64_t)p.chunk_bytes, (uint32_t)p.chunk_bytes, &full[s], pol);
                bulk_g2s(st + p.chunk_bytes, p.xq + (c % p.KC) * (int64_t)xbytes, xbytes, &full[s]);
            }
        }
        return;
    }

    // ===================== consumers =====================
    const int g4 = lane >> 3, i = lane & 7;
    const FmtMeta meta{p.gpc};
    float acc[4][MB];
#pragma unroll
    for (int s4 = 0; s4 < 4; s4++)
#pragma unroll
        for (int m = 0; m < MB; m++) acc[s4][m] = 0.0f;

    int64_t it = 0;
    for (int64_t c = c0; c < c1; c++, it++) {
        const int s = (int)(it % nst);
        const uint32_t round = (uint32_t)(it / nst);
        mbar_wait(&full[s], round & 1u);
        const uint8_t* wc = stages + (size_t)s * p.stage_bytes;
        const uint8_t* xr = wc + p.chunk_bytes;

        uint4 xa[MB], xb[MB];
        float dx[MB];
        int bsA[MB], bsB[MB];
#pragma unroll
        for (int m = 0; m < MB; m++) {
            const uint8_t* rec = xr + m * ACT_REC_BYTES;
            xa[m] = lds128(rec + 32 * i);
            xb[m] = lds128(rec + 32 * i + 16);
            dx[m] = *reinterpret_cast<const float*>(rec + 256 + 4 * i);
            uint32_t bs = *reinterpret_cast<const uint32_t*>(rec + 288 + 4 * i);
            bsA[m] = (int)(int16_t)(bs & 0xFFFFu);
            bsB[m] = (int)(int16_t)(bs >> 16);
        }
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
            const int r = 16 * warp + 4 * s4 + g4;
            Unit u;
            F::template load_unit<true>(wc, r, i, u, meta);
#pragma unroll
            for (int m = 0; m < MB; m++) {
                int sA = 0, sB = 0;
                sA = __dp4a((int)u.v[0], (int)xa[m].x, sA); sA = __dp4a((int)u.v[1], (int)xa[m].y, sA);
                sA = __dp4a((int)u.v[2], (int)xa[m].z, sA); sA = __dp4a((int)u.v[3], (int)xa[m].w, sA);
                sB = __dp4a((int)u.v[4], (int)xb[m].x, sB); sB = __dp4a((int)u.v[5], (int)xb[m].y, sB);
                sB = __dp4a((int)u.v[6], (int)xb[m].z, sB); sB = __dp4a((int)u.v[7], (int)xb[m].w, sB);
                sA -= u.off[0] * bsA[m];
                sB -= u.off[1] * bsB[m];
                float t = (u.a[0] * dx[m]) * (float)sA + (u.a[1] * dx[m]) * (float)sB;
                if (F::HAS_MIN) t -= (u.b[0] * dx[m]) * (float)bsA[m] + (u.b[1] * dx[m]) * (float)bsB[m];
                acc[s4][m] += t;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);

        // ---- tile boundary: flush ----
        const bool tile_end = ((c + 1) % p.KC == 0) || (c + 1 == c1);
        if (tile_end) {
            const int64_t t = c / p.KC;
#pragma unroll
            for (int s4 = 0; s4 < 4; s4++)
#pragma unroll
                for (int m = 0; m < MB; m++) {
                    float v = acc[s4][m];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    acc[s4][m] = v;
                }
            const bool full_tile = (c0 <= t * p.KC) && ((t + 1) * p.KC <= c1);
            if (full_tile) {
                if (i == 0) {
#pragma unroll
                    for (int s4 = 0; s4 < 4; s4++) {
                        const int64_t n = t * TILE_ROWS + 16 * warp + 4 * s4 + g4;
                        if (n < p.N) {
                            const float bv = p.bias ? p.bias[n] : 0.0f;
#pragma unroll
                            for (int m = 0; m < MB; m++)
                                if (m < p.M) store_out(p.y, p.y_dtype, (int64_t)m * p.ldy + n, acc[s4][m] + bv);
                        }
                    }
                }
            } else {
                const int slot = (t == c0 / p.KC) ? 0 : 1;
                float* part = p.ws_part + ((size_t)g * 2 + slot) * (TILE_ROWS * MB);
                if (i == 0) {
#pragma unroll
                    for (int s4 = 0; s4 < 4; s4++) {
                        const int rr = 16 * warp + 4 * s4 + g4;
#pragma unroll
                        for (int m = 0; m < MB; m++) part[rr * MB + m] = acc[s4][m];
                    }
                    __threadfence();
                }
                named_bar_sync(1, MV_CONSUMER_WARPS * 32);
                const int64_t gf = sk_owner(t * p.KC, p.C, G), gl = sk_owner((t + 1) * p.KC - 1, p.C, G);
                if (tid == 0) {
                    __threadfence();
                    unsigned int old = atomicAdd(&p.ws_cnt[t], 1u);
                    *sflag = (old == (unsigned int)(gl - gf)) ? 1 : 0;
                }
                named_bar_sync(1, MV_CONSUMER_WARPS * 32);
                if (*sflag) {
                    __threadfence();
                    for (int idx = tid; idx < TILE_ROWS * MB; idx += MV_CONSUMER_WARPS * 32) {
                        float sum = 0.0f;
                        for (int64_t gg = gf; gg <= gl; gg++) {
                            const int sl = (sk_begin(gg, p.C, G) / p.KC == t) ? 0 : 1;
                            sum += __ldcg(p.ws_part + ((size_t)gg * 2 + sl) * (TILE_ROWS * MB) + idx);
                        }
                        const int rr = idx / MB, m = idx % MB;
                        const int64_t n = t * TILE_ROWS + rr;
                        if (n < p.N && m < p.M) store_out(p.y, p.y_dtype, (int64_t)m * p.ldy + n, sum + (p.bias ? p.bias[n] : 0.0f));
                    }
                    if (tid == 0) p.ws_cnt[t] = 0u;
                }
                named_bar_sync(1, MV_CONSUMER_WARPS * 32);
            }
#pragma unroll
            for (int s4 = 0; s4 < 4; s4++)
#pragma unroll
                for (int m = 0; m < MB; m++) acc[s4][m] = 0.0f;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <class F, int MB>
static cudaError_t launch_t(const MatvecParams& p, int grid, int smem, cudaStream_t st) {
    static bool configured[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(matvec_kernel<F, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    matvec_kernel<F, MB><<<grid, MV_THREADS, smem, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

template <class F>
static cudaError_t launch_f(const MatvecParams& p, int mb, int grid, int smem, cudaStream_t st) {
    switch (mb) {
        case 1: return launch_t<F, 1>(p, grid, smem, st);
        case 2: return launch_t<F, 2>(p, grid, smem, st);
        case 4: return launch_t<F, 4>(p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t matvec_plan(const b200q_weight* w, int64_t M, MatvecPlan* plan) {
    if (M < 1 || M > 4) return cudaErrorInvalidValue;
    int mb = M == 1 ? 1 : (M == 2 ? 2 : 4);
    int stage = w->chunk_bytes + (int)M * ACT_REC_BYTES;
    stage = (stage + 127) & ~127;
    int budget = 108 * 1024 - MV_HDR_BYTES;
    int nst = budget / stage;
    if (nst > MV_MAX_STAGES) nst = MV_MAX_STAGES;
    if (nst < 2) return cudaErrorInvalidValue;
    int64_t C = w->T * w->KC;
    int ctas_per_sm = (mb <= 2) ? 2 : 1;
    int64_t G = (int64_t)w->num_sms * ctas_per_sm;
    if (G > C) G = C;
    plan->grid = (int)G;
    plan->nstages = nst;
    plan->stage_bytes = stage;
    plan->smem_bytes = MV_HDR_BYTES + nst * stage;
    plan->mb = mb;
    return cudaSuccess;
}

size_t matvec_ws_bytes(const b200q_weight* w, int64_t M) {
    // counters (one per row tile, padded) + 2 partial slots per potential CTA
    size_t cnt = ((size_t)w->T * 4 + 255) & ~(size_t)255;
    size_t part = (size_t)w->num_sms * 2 * 2 * TILE_ROWS * 4 * sizeof(float);
    (void)M;
    return cnt + part;
}

cudaError_t launch_matvec(const b200q_weight* w, const uint8_t* xq, int64_t M, void* y, int y_dtype, int64_t ldy, uint8_t* ws, cudaStream_t st) {
    MatvecPlan plan;
    cudaError_t e = matvec_plan(w, M, &plan);
    if (e != cudaSuccess) return e;
    MatvecParams p;
    p.w = w->data;
    p.xq = xq;
    p.y = y;
    p.bias = w->bias;
    size_t cnt = ((size_t)w->T * 4 + 255) & ~(size_t)255;
    p.ws_cnt = reinterpret_cast<unsigned int*>(ws);
    p.ws_part = reinterpret_cast<float*>(ws + cnt);
    p.N = w->N;
    p.M = (int)M;
    p.y_dtype = y_dtype;
    p.ldy = ldy;
    p.KC = w->KC;
    p.C = w->T * w->KC;
    p.gpc = w->gpc;
    p.nstages = plan.nstages;
    p.chunk_bytes = w->chunk_bytes;
    p.stage_bytes = plan.stage_bytes;
    switch (w->family) {
        case B200Q_FAM_Q4_K: return launch_f<FmtQ4K>(p, plan.mb, plan.grid, plan.smem_bytes, st);
        case B200Q_FAM_Q6_K: return launch_f<FmtQ6K>(p, plan.mb, plan.grid, plan.smem_bytes, st);
        case B200Q_FAM_Q8_0: return launch_f<FmtQ8_0>(p, plan.mb, plan.grid, plan.smem_bytes, st);
        case B200Q_FAM_G4: return launch_f<FmtG4>(p, plan.mb, plan.grid, plan.smem_bytes, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace b200q
