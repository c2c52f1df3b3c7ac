// gemm_tc.cu -- dequant-GEMM on the 5th-generation tensor cores (tcgen05 + TMEM) for M >= 5:
// batched decode (HBM-bound) and prefill (tensor-bound) through one kernel.
//
// Design (DESIGN.md "Kernel 2"): Y^T tile = W_tile[128 n x K] . X_tile[Mt m x K]^T
//   * UMMA M = 128 weight rows, UMMA N = Mt <= 256 activation rows, K = 16 per instruction, f16 x f16 -> f32;
//   * the weight operand never touches shared memory as f16: packed chunks arrive by cp.async.bulk (TMA
//     engine), 8 dequant warps (thread == weight row, bank-conflict-free thanks to the upload swizzle) expand
//     them in registers and write f16 pairs straight into TENSOR MEMORY with tcgen05.st; the MMA reads A
//     from TMEM (the ".ts" form), so shared-memory bandwidth is spent only on the 4.5-8.5 bit packed bytes
//     and on the activation tile;
//   * the activation tile comes from a pre-staged f16 copy laid out as 128B-swizzled K-major UMMA tiles, one
//     contiguous cp.async.bulk per 64-k sub-stage (no tensor map needed);
//   * accumulators (128 lanes x Mt columns f32) live in TMEM; 8 TMEM A-slots of 64 k form the
//     dequant -> MMA ring; tcgen05.commit releases A-slots / X stages and publishes the accumulator;
//   * warp roles: 0 = weight-chunk producer, 2 = activation producer, 1 = TMEM allocator + single-thread
//     MMA issuer, 4..11 = dequant (8 warps) + epilogue (tcgen05.ld -> bias -> global).
#pragma once
#include <atomic>
#include "formats.cuh"
#include "internal.h"

namespace b200q {


// dequant warps (template parameter DQW): DQW/4 per TMEM lane quarter, each 4/DQW of the chunk's k.  8 for wide M tiles
// (the MMA is the bottleneck; 16 measured slower there), 16 for skinny M (<= 64), where every weight is expanded for a
// handful of MMAs and the expansion's latency chains are what bounds the kernel (ncu r1: 2 warps/scheduler, IPC 0.4).
constexpr int GT_NX = 16;      // max activation sub-stage ring depth (64 k each); runtime depth p.nx
constexpr int GT_ASLOTS = 3;   // max TMEM A slots; one slot = one whole 256-k chunk of f16 (128 columns)
constexpr int GT_MAX_NW = 4;   // weight chunk ring
constexpr int GT_D_COL = 0;    // accumulator columns [0, 256)
// A slots occupy the top of TMEM: columns [512 - 128*nslots, 512); 2 slots when Mt > 128, else 3
constexpr int GT_HDR = 1024;

struct GemmParams {
    const uint8_t* w;
    const uint8_t* xs;
    void* y;
    const float* bias;
    int64_t N, M, ldy;
    int y_dtype;
    int T, KC, MT, Mt;
    int gpc, chunk_bytes, nw, w_stage_bytes, x_stage_bytes;
    int nslots, a_col, nx, xsub;   // xsub: 64-k sub-tiles per X stage (1 or 4)
    int dq_warps;      // 8 (wide M tiles) or 16 + dedicated epilogue warps + two accumulators (Mt <= 128)
    int splits;        // split-K factor (serial K loop is the latency floor when there are fewer tiles than SMs)
    float* partial;    // [splits][M][N] f32 when splits > 1
    uint32_t idesc;
    // tail split-K (wide M): tiles [0, tail_first) are whole work items; the tail_n tiles of the last, partial wave are cut
    // tail_splits ways along K so that wave is short and full (896 tiles on 148 SMs otherwise run 7 waves for 6.05 waves of work);
    // their partial accumulators go to partial_tail [tail_n][tail_splits][128][Mt] f32 and splitk_tail_reduce_kernel sums them
    int tail_first, tail_n, tail_splits;
    float* partial_tail;
    int xmc;           // 1: launched as clusters of two CTAs that work on ADJACENT weight tiles of the same activation tile; each
                       // CTA fetches half of every activation stage and TMA-multicasts it to both (halves the L2 -> SM traffic
                       // that bounds the wide-M kernel: 128 KB of X per 27 KB chunk of W)
};

// work item -> (weight tile t, activation tile mt, K split sp).  Plain: items = tiles, one CTA each.  xmc: items are tile
// PAIRS (2 tp, 2 tp + 1) walked by a cluster; the odd CTA of a trailing pair recomputes tile T-1 and stores nothing.
struct GemmItem {
    int t, mt, sp, nsp;   // weight tile, activation tile, K split index / count of this item
    int kc0, kc1;         // k-chunk range
    int tail;             // >= 0: index of this tile among the tail tiles (partials go to partial_tail); -1 otherwise
    bool valid;
};
__device__ __forceinline__ int gemm_first(const GemmParams& p) { return p.xmc ? (int)(blockIdx.x >> 1) : (int)blockIdx.x; }
__device__ __forceinline__ int gemm_stride(const GemmParams& p) { return p.xmc ? (int)(gridDim.x >> 1) : (int)gridDim.x; }
__device__ __forceinline__ int gemm_total(const GemmParams& p) {
    if (p.tail_splits > 1) return p.tail_first + p.tail_n * p.tail_splits;
    return (p.xmc ? (p.T + 1) / 2 : p.T) * p.MT * p.splits;
}
__device__ __forceinline__ GemmItem gemm_item(const GemmParams& p, int it) {
    GemmItem g;
    int rest;
    if (p.tail_splits > 1) {   // (splits == 1, no clusters)
        if (it >= p.tail_first) {
            const int r = it - p.tail_first;
            g.tail = r / p.tail_splits;
            g.sp = r % p.tail_splits;
            g.nsp = p.tail_splits;
            rest = p.tail_first + g.tail;
        } else {
            g.tail = -1; g.sp = 0; g.nsp = 1;
            rest = it;
        }
    } else {
        g.tail = -1;
        g.sp = it % p.splits;
        g.nsp = p.splits;
        rest = it / p.splits;
    }
    g.mt = rest % p.MT;
    int t = rest / p.MT;
    if (p.xmc) t = 2 * t + (int)(blockIdx.x & 1);
    g.valid = t < p.T;
    g.t = g.valid ? t : p.T - 1;
    g.kc0 = g.sp * p.KC / g.nsp;
    g.kc1 = (g.sp + 1) * p.KC / g.nsp;
    return g;
}
// where one accumulator element of a split item goes
__device__ __forceinline__ float* gemm_partial_ptr(const GemmParams& p, const GemmItem& g, int r, int64_t n, int64_t m, int m_local) {
    if (g.tail >= 0) return p.partial_tail + (((size_t)g.tail * p.tail_splits + g.sp) * TILE_ROWS + r) * p.Mt + m_local;
    return p.partial + ((size_t)g.sp * p.M + m) * p.N + n;
}

// ---- tcgen05 wrappers ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t cta_mask) {  // arrive on the same barrier of every CTA in the mask
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory operand descriptor (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;              // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset
    d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;              // SWIZZLE_128B
    return d;
}

// 32 integer weights of a unit (one byte each) -> 16 packed f16 pairs  w = a*(v-off) - b,  rounded once to f16.
// Bytes become f16 with the 0x6400 trick (0x6400 | v == 1024 + v exactly), the integer offset is removed with an
// exact HSUB2 and the scale/min applied with one HFMA2: one PRMT + HSUB2 + HFMA2 per pair of weights, no I2F.
// Pair order: word k yields (e0,e2) then (e1,e3) -- the activation staging kernel applies the same [0,2,1,3]
// permutation inside every group of 4 k, so the contraction is unchanged.
template <bool SIGNED>
__device__ __forceinline__ void unit_to_f16(const Unit& u, uint32_t* out) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const float offf = 1024.0f + (float)u.off[h] + (SIGNED ? 128.0f : 0.0f);
        const __half2 off2 = __float2half2_rn(offf);         // integer <= 2047: exact in f16
        const __half2 a2 = __float2half2_rn(u.a[h]);
        const __half2 nb2 = __float2half2_rn(-u.b[h]);
#pragma unroll
        for (int k = 4 * h; k < 4 * h + 4; k++) {
            const uint32_t wv = SIGNED ? (u.v[k] ^ 0x80808080u) : u.v[k];
            uint32_t p02 = __byte_perm(wv, 0x64646464u, 0x4240);      // bytes [e0, 0x64, e2, 0x64] = (1024 + e0, 1024 + e2): one PRMT
            uint32_t p13 = __byte_perm(wv, 0x64646464u, 0x4341);      // (1024 + e1, 1024 + e3)
            __half2 x02 = __hsub2(*reinterpret_cast<__half2*>(&p02), off2);
            __half2 x13 = __hsub2(*reinterpret_cast<__half2*>(&p13), off2);
            x02 = __hfma2(x02, a2, nb2);
            x13 = __hfma2(x13, a2, nb2);
            out[2 * k] = *reinterpret_cast<uint32_t*>(&x02);
            out[2 * k + 1] = *reinterpret_cast<uint32_t*>(&x13);
        }
    }
}

// EPI: four dedicated epilogue warps (one per TMEM lane quarter) and two accumulator buffers (Mt <= 128), so the
// dequant warps start the next work item while the previous accumulator drains -- skinny launches are made of many
// short items and the drain (MMA latency + tcgen05.ld + global stores) otherwise sits between every pair of them.
template <class F, int DQW, bool EPI>
__global__ void __launch_bounds__((4 + DQW + (EPI ? 4 : 0)) * 32, 1) gemm_tc_kernel(const GemmParams p) {
    constexpr int GT_DQ_WARPS = DQW;
    constexpr int KG = DQW / 4;        // k groups: warp (q, h) expands units KG*j + h, j < 8 / KG
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_w = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_w = full_w + GT_MAX_NW;
    uint64_t* full_x = empty_w + GT_MAX_NW;
    uint64_t* empty_x = full_x + GT_NX;
    uint64_t* a_full = empty_x + GT_NX;
    uint64_t* a_empty = a_full + GT_ASLOTS;
    uint64_t* d_full = a_empty + GT_ASLOTS;   // [2]
    uint64_t* d_empty = d_full + 2;           // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d_empty + 2);
    uint8_t* xst = smem + GT_HDR;
    uint8_t* wst = xst + (size_t)p.nx * p.x_stage_bytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < p.nw; s++) { mbar_init(&full_w[s], 1); mbar_init(&empty_w[s], GT_DQ_WARPS); }
        for (int s = 0; s < p.nx; s++) { mbar_init(&full_x[s], 1); mbar_init(&empty_x[s], p.xmc ? 2 : 1); }  // xmc: both CTAs' MMAs release a stage
        for (int s = 0; s < p.nslots; s++) { mbar_init(&a_full[s], GT_DQ_WARPS); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], 4); }
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (p.xmc) cluster_sync_all();  // the peer's barriers exist before anything is multicast into this CTA
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    // PDL: the weight stream (static data) starts right away and overlaps the predecessor's tail; only the activation
    // producer (reads the staged tiles) and the epilogues (write y / the split-K partials) wait for the preceding kernels
    pdl_launch_dependents();

    const int total_tiles = gemm_total(p);
    const int it0 = gemm_first(p), its = gemm_stride(p);
    const int KS = p.KC * 4;  // 64-k sub-stages along K

    if (warp == 0) {
        // ===================== weight chunk producer =====================
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            int s = 0;
            uint32_t ph = 1;
            for (int tile = it0; tile < total_tiles; tile += its) {
                const GemmItem gi = gemm_item(p, tile);
                const int t = gi.t;
                const int kc0 = gi.kc0, kc1 = gi.kc1;
                const uint8_t* src = p.w + ((size_t)t * p.KC + kc0) * p.chunk_bytes;
                for (int kc = kc0; kc < kc1; kc++) {
                    mbar_wait(&empty_w[s], ph);
                    mbar_arrive_expect_tx(&full_w[s], (uint32_t)p.chunk_bytes);
                    if (p.MT == 1) bulk_g2s_hint(wst + (size_t)s * p.w_stage_bytes, src, (uint32_t)p.chunk_bytes, &full_w[s], pol);
                    else bulk_g2s(wst + (size_t)s * p.w_stage_bytes, src, (uint32_t)p.chunk_bytes, &full_w[s]);
                    src += p.chunk_bytes;
                    if (++s == p.nw) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== activation tile producer =====================
        if (lane == 0) {
            pdl_wait();
            int s = 0;
            uint32_t ph = 1;
            const uint32_t half = (uint32_t)p.x_stage_bytes >> 1, hoff = (uint32_t)(blockIdx.x & 1) * half;
            for (int tile = it0; tile < total_tiles; tile += its) {
                const GemmItem gi = gemm_item(p, tile);
                const int mt = gi.mt;
                const int kc0 = gi.kc0, kc1 = gi.kc1;
                // one X stage = xsub consecutive 64-k sub-tiles (contiguous in the staged layout): 1 for wide M tiles,
                // 4 (a whole chunk) for skinny ones, where per-stage handshakes would otherwise pace the MMA issuer
                const uint8_t* src = p.xs + ((size_t)mt * KS + 4 * kc0) * (size_t)(p.Mt * 128);
                for (int ks = 4 * kc0; ks < 4 * kc1; ks += p.xsub) {
                    mbar_wait(&empty_x[s], ph);   // xmc: BOTH CTAs have consumed this stage (their commits are multicast)
                    mbar_arrive_expect_tx(&full_x[s], (uint32_t)p.x_stage_bytes);
                    if (p.xmc) bulk_g2s_mc(xst + (size_t)s * p.x_stage_bytes + hoff, src + hoff, half, &full_x[s], (uint16_t)3);  // my half -> both CTAs
                    else bulk_g2s(xst + (size_t)s * p.x_stage_bytes, src, (uint32_t)p.x_stage_bytes, &full_x[s]);
                    src += p.x_stage_bytes;
                    if (++s == p.nx) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== single-thread MMA issuer =====================
        if (elect_one()) {
            int xs = 0, as = 0, acc = 0;
            uint32_t xph = 0, aph = 0, eph0 = 1, eph1 = 1;
            for (int tile = it0; tile < total_tiles; tile += its) {
                const GemmItem gi = gemm_item(p, tile);
                const int kc0 = gi.kc0, kc1 = gi.kc1;
                if (EPI) {  // the epilogue warps have drained this accumulator buffer (two items ago)
                    mbar_wait(&d_empty[acc], acc ? eph1 : eph0);
                    if (acc) eph1 ^= 1u; else eph0 ^= 1u;
                    tc_fence_after();
                }
                const uint32_t d_col = GT_D_COL + (EPI ? acc * p.Mt : 0);
                for (int kc = kc0; kc < kc1; kc++) {
                    mbar_wait(&a_full[as], aph);  // a whole dequantised chunk (4 x 64 k) is in TMEM
                    const uint32_t a_chunk = tmem + p.a_col + as * 128;
                    if (p.xsub == 4) {
                        mbar_wait(&full_x[xs], xph);
                        tc_fence_after();
                        const uint64_t bdesc0 = make_b_desc(smem_u32(xst + (size_t)xs * p.x_stage_bytes));
                        const uint64_t sub = (uint64_t)((p.Mt * 128) >> 4);  // descriptor address units of 16 B
#pragma unroll
                        for (int j = 0; j < 4; j++)
#pragma unroll
                            for (int kk = 0; kk < 4; kk++)
                                tc_mma_ts(tmem + d_col, a_chunk + j * 32 + kk * 8, bdesc0 + j * sub + (uint64_t)(kk * 2), p.idesc, ((kc - kc0) | j | kk) != 0 ? 1u : 0u);
                        if (p.xmc) tc_commit_mc(&empty_x[xs], (uint16_t)3); else tc_commit(&empty_x[xs]);
                        if (++xs == p.nx) { xs = 0; xph ^= 1u; }
                    } else {
#pragma unroll 1
                        for (int j = 0; j < 4; j++) {
                            mbar_wait(&full_x[xs], xph);
                            tc_fence_after();
                            const uint64_t bdesc = make_b_desc(smem_u32(xst + (size_t)xs * p.x_stage_bytes));
#pragma unroll
                            for (int kk = 0; kk < 4; kk++)
                                tc_mma_ts(tmem + d_col, a_chunk + j * 32 + kk * 8, bdesc + (uint64_t)(kk * 2), p.idesc, ((kc - kc0) | j | kk) != 0 ? 1u : 0u);
                            if (p.xmc) tc_commit_mc(&empty_x[xs], (uint16_t)3); else tc_commit(&empty_x[xs]);
                            if (++xs == p.nx) { xs = 0; xph ^= 1u; }
                        }
                    }
                    tc_commit(&a_empty[as]);
                    if (++as == p.nslots) { as = 0; aph ^= 1u; }
                }
                tc_commit(&d_full[acc]);
                if (EPI) acc ^= 1;
            }
        }
    } else if (EPI && warp >= 4 + DQW) {
        // ===================== dedicated epilogue warps: accumulator rows = n, columns = m =====================
        const int q = warp & 3;
        const int r = 32 * q + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
        int acc = 0;
        uint32_t ph0 = 0, ph1 = 0;
        pdl_wait();  // y / partial may still be read by the preceding kernels
        for (int tile = it0; tile < total_tiles; tile += its) {
            const GemmItem gi = gemm_item(p, tile);
            const int t = gi.t, mt = gi.mt;
            mbar_wait(&d_full[acc], acc ? ph1 : ph0);
            if (acc) ph1 ^= 1u; else ph0 ^= 1u;
            tc_fence_after();
            const int64_t n = (int64_t)t * TILE_ROWS + r;
            const float bv = (p.bias && n < p.N) ? p.bias[n] : 0.0f;
            for (int cb = 0; cb < p.Mt; cb += 16) {
                uint32_t v[16];
                tc_ld16(lane_base + GT_D_COL + acc * p.Mt + cb, v);
                tc_wait_ld();
                if (cb + 16 >= p.Mt) {  // last read of this buffer: hand it back to the MMA issuer before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d_empty[acc]);
                }
                if (n < p.N && gi.valid) {
#pragma unroll
                    for (int c = 0; c < 16; c++) {
                        const int64_t m = (int64_t)mt * p.Mt + cb + c;
                        if (m < p.M) {
                            if (gi.nsp > 1) *gemm_partial_ptr(p, gi, r, n, m, cb + c) = __uint_as_float(v[c]);
                            else store_out(p.y, p.y_dtype, m * p.ldy + n, __uint_as_float(v[c]) + bv);
                        }
                    }
                }
            }
            acc ^= 1;
        }
    } else if (warp >= 4) {
        // ===================== dequant (thread == weight row) + epilogue =====================
        const int q = warp & 3, h = (warp - 4) >> 2;  // lane quarter; k group (units KG*j+h); h < 2 also = epilogue column half
        const int r = 32 * q + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
        const FmtMeta meta{p.gpc};
        int ws = 0, as = 0;
        uint32_t wph = 0, aph = 1, dph = 0;
        if (!EPI) pdl_wait();  // these warps also run the epilogue (stores to y / partial)
        for (int tile = it0; tile < total_tiles; tile += its) {
            const GemmItem gi = gemm_item(p, tile);
            const int t = gi.t, mt = gi.mt;
            const int kc0 = gi.kc0, kc1 = gi.kc1;
            for (int kc = kc0; kc < kc1; kc++) {
                mbar_wait(&full_w[ws], wph);
                const uint8_t* wc = wst + (size_t)ws * p.w_stage_bytes;
                mbar_wait(&a_empty[as], aph);  // the MMAs that read this TMEM slot two/three chunks ago are done
                tc_fence_after();
#pragma unroll 1
                for (int j = 0; j < 8 / KG; j++) {
                    Unit u;
                    F::template load_unit<true>(wc, r, KG * j + h, u, meta);
                    uint32_t pk[16];
                    unit_to_f16<F::SIGNED>(u, pk);
                    tc_st16(lane_base + p.a_col + as * 128 + (KG * j + h) * 16, pk);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[as]);
                if (++as == p.nslots) { as = 0; aph ^= 1u; }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_w[ws]);
                if (++ws == p.nw) { ws = 0; wph ^= 1u; }
            }
            if (EPI) continue;  // the dedicated epilogue warps drain the accumulator
            // ---- epilogue: accumulator rows = n, columns = m ----
            mbar_wait(&d_full[0], dph);
            dph ^= 1u;
            tc_fence_after();
            const int64_t n = (int64_t)t * TILE_ROWS + r;
            const float bv = (p.bias && n < p.N) ? p.bias[n] : 0.0f;
            const int half_cols = p.Mt >> 1;
            for (int cb = 0; cb < (h < 2 ? half_cols : 0); cb += 16) {
                uint32_t v[16];
                tc_ld16(lane_base + GT_D_COL + h * half_cols + cb, v);
                tc_wait_ld();
                if (n < p.N && gi.valid) {
#pragma unroll
                    for (int c = 0; c < 16; c++) {
                        const int64_t m = (int64_t)mt * p.Mt + h * half_cols + cb + c;
                        if (m < p.M) {
                            if (gi.nsp > 1) *gemm_partial_ptr(p, gi, r, n, m, h * half_cols + cb + c) = __uint_as_float(v[c]);
                            else store_out(p.y, p.y_dtype, m * p.ldy + n, __uint_as_float(v[c]) + bv);
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (p.xmc) cluster_sync_all();  // the peer may still multicast into this CTA's stages / arrive on its barriers until it is done too
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

template <class F, int DQW, bool EPI>
static cudaError_t launch_gemm_dq(const GemmParams& p, int grid, int smem, cudaStream_t st) {
    static std::atomic<bool> configured[16];   // per device; racing first calls both set the attribute (idempotent): re-entrant
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 16 || !configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<F, DQW, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        if (dev < 16) configured[dev].store(true, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((4 + DQW + (EPI ? 4 : 0)) * 32);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: see the kernel's griddepcontrol placement
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
    if (p.xmc) {  // clusters of two CTAs (adjacent weight tiles, shared activation stages)
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 2;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    cfg.attrs = attr;
    cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<F, DQW, EPI>, p);
    count_launch();
    return le != cudaSuccess ? le : cudaGetLastError();
}
template <class F>
static cudaError_t launch_gemm_t(const GemmParams& p, int grid, int smem, cudaStream_t st) {
    return p.dq_warps == 16 ? launch_gemm_dq<F, 16, true>(p, grid, smem, st) : launch_gemm_dq<F, 8, false>(p, grid, smem, st);
}


// per-format launcher, explicitly specialised in inst_<format>.cu
template <int FAMILY>
cudaError_t gemm_launch(const GemmParams& p, int grid, int smem, cudaStream_t st);

}  // namespace b200q
