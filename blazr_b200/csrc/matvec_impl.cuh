// matvec.cu -- decode-time dequant-matvec (M = 1..4): the HBM-bound hot kernel.
//
// Design (DESIGN.md "Kernel 1"):
//   * stream-K over the tile-major chunk array: CTA g owns the contiguous chunk range
//     [g*C/G, (g+1)*C/G)  -> every SM streams one contiguous region of HBM, perfect balance for any N,K;
//   * one producer thread feeds a ring of shared-memory stages with cp.async.bulk (TMA engine, 1-D,
//     mbarrier complete_tx): 1 copy of the 18-35 KB weight chunk + 1 copy of the M x 320 B quantised
//     activation record per stage, so ~100 KB per SM are in flight independent of occupancy;
//   * 16 consumer warps: warp w owns rows 8w..8w+7 of the tile, 8 lanes x 16 B walk one row's
//     256-k chunk (unit i = 32 consecutive k), 4 rows per step; activations stay in registers for the
//     8 rows; integer dot products with dp4a, per-sub-block scales applied in f32;
//   * at tile end an 8-lane shuffle reduction; tiles split between CTAs are combined DETERMINISTICALLY
//     through per-CTA partial slots in the workspace: the last CTA to arrive (atomic counter) sums the
//     partials in CTA order and writes y (no float atomics, no inter-CTA waiting).
#pragma once
#include "matvec_common.cuh"

namespace b200q {

// ------------------------------------------------------------------------------------------------
// Fused activation producers.  Executed by the 16 consumer warps after griddepcontrol.wait, while the first
// ring-full of weight chunks (requested before the wait) is still in flight, so the separate norm / SwiGLU
// kernels of a decode step (and their launch + drain gaps) disappear.  Arithmetic is identical to
// decode_ops.cu (f64 sum of squares, explicit f32 ops, deterministic exp) => same bits as the oracle.
// Records: xhat[(kc*M + m)*320] = { int8 q[256]; float d[8]; (int16 bsum16[2])[8] }.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void quant_block_to_record(float v, uint8_t* rec, int blk_in_chunk, int lane) {
    float amax = fabsf(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float d = __fdiv_rn(amax, 127.0f);
    const float id = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
    const int q = (int)roundf(__fmul_rn(v, id));
    int s = q;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int s_hi = __shfl_sync(0xffffffffu, s, 16);
    rec[blk_in_chunk * 32 + lane] = (uint8_t)(int8_t)q;
    if (lane == 0) {
        reinterpret_cast<float*>(rec + 256)[blk_in_chunk] = d;
        reinterpret_cast<uint32_t*>(rec + 288)[blk_in_chunk] = ((uint32_t)s & 0xFFFFu) | ((uint32_t)s_hi << 16);
    }
}

template <int MB>
__device__ __forceinline__ void fused_prologue(const MatvecParams& p, uint8_t* xhat, int tid, int warp, int lane, bool writer) {
    constexpr int NT = MV_CONSUMER_WARPS * 32;
    const int K = (int)p.KC * CHUNK_K;
    const int nblk = K / 32;
    if (p.pro == 1) {
        // ---- h = h_in (+ delta); xhat = quant(rmsnorm(h) * w) ----
        // warp w owns the 32-blocks w, w+16, ...; every element is loaded once (all loads in flight together),
        // kept in registers for the sum of squares and then normalised + quantised from registers.
        constexpr int VMAX = 16;  // K <= 8192
        __shared__ double red[MB][MV_CONSUMER_WARPS];
        __shared__ float s_inv[MB];
        const int nv = (nblk + MV_CONSUMER_WARPS - 1) / MV_CONSUMER_WARPS;
        float v[MB][VMAX];
#pragma unroll
        for (int m = 0; m < MB; m++) {
            if (m >= p.M) break;
            const float* hr = p.h_in + (size_t)m * K;
            const float* dr = p.delta ? p.delta + (size_t)m * K : nullptr;
#pragma unroll
            for (int j = 0; j < VMAX; j++) {
                const int b = warp + j * MV_CONSUMER_WARPS;
                v[m][j] = (j < nv && b < nblk) ? hr[b * 32 + lane] : 0.0f;
            }
            if (dr) {
#pragma unroll
                for (int j = 0; j < VMAX; j++) {
                    const int b = warp + j * MV_CONSUMER_WARPS;
                    if (j < nv && b < nblk) v[m][j] = __fadd_rn(v[m][j], dr[b * 32 + lane]);
                }
            }
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int j = 0; j < VMAX; j += 2) {
                s0 = fma((double)v[m][j], (double)v[m][j], s0);
                s1 = fma((double)v[m][j + 1], (double)v[m][j + 1], s1);
            }
            double ss = s0 + s1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) red[m][warp] = ss;
            if (writer && p.h_out) {
#pragma unroll
                for (int j = 0; j < VMAX; j++) {
                    const int b = warp + j * MV_CONSUMER_WARPS;
                    if (j < nv && b < nblk) p.h_out[(size_t)m * K + b * 32 + lane] = v[m][j];
                }
            }
        }
        named_bar_sync(4, NT);
        if (tid < p.M) {
            double tot = 0.0;
#pragma unroll
            for (int i = 0; i < MV_CONSUMER_WARPS; i++) tot += red[tid][i];
            s_inv[tid] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn((float)tot, (float)K), p.eps)));
        }
        named_bar_sync(4, NT);
#pragma unroll
        for (int m = 0; m < MB; m++) {
            if (m >= p.M) break;
            const float inv = s_inv[m];
#pragma unroll
            for (int j = 0; j < VMAX; j++) {
                const int b = warp + j * MV_CONSUMER_WARPS;
                if (j < nv && b < nblk) {
                    const float x = __fmul_rn(__fmul_rn(v[m][j], inv), p.norm_w[b * 32 + lane]);
                    quant_block_to_record(x, xhat + ((size_t)(b >> 3) * p.M + m) * ACT_REC_BYTES, b & 7, lane);
                }
            }
        }
    } else {
        // ---- xhat = quant(silu(gate) * up), gate_up[M, 2K] ----
        for (int m = 0; m < p.M; m++) {
            const float* gr = p.gate_up + (size_t)m * 2 * K;
            for (int b = warp; b < nblk; b += MV_CONSUMER_WARPS) {
                const int k = b * 32 + lane;
                const float gv = gr[k], uv = gr[K + k];
                const float v = __fmul_rn(__fdiv_rn(gv, __fadd_rn(1.0f, det_expf(-gv))), uv);
                quant_block_to_record(v, xhat + ((size_t)(b >> 3) * p.M + m) * ACT_REC_BYTES, b & 7, lane);
            }
        }
    }
    named_bar_sync(4, NT);  // records visible to every consumer warp
}

// ---- output of one finished row sum: local y, or (fused TP exchange) the same element of every rank's slot ----
struct RpState {
    unsigned int epoch;  // epoch of this exchange (previous + 1)
    size_t off;          // byte offset of this rank's slot / gather region inside every rank's buffer
};
__device__ __forceinline__ RpState rp_begin(const MatvecParams& p) {
    RpState r{0u, 0};
    if (p.rp_mode == RP_NONE) return r;
    const uint8_t* mine = p.comm.peers[p.comm.rank];
    const bool ar = p.rp_mode == RP_ALLREDUCE;
    // written by the last CTA of the previous exchange of this kind; that launch completed before griddepcontrol.wait returned
    r.epoch = __ldcg(reinterpret_cast<const unsigned int*>(mine + (ar ? COMM_OFF_AR_EPOCH : COMM_OFF_AG_EPOCH))) + 1u;
    r.off = ar ? comm_ar_slot_off(p.comm, (int)(r.epoch & 1u), p.comm.rank) : comm_ag_off(p.comm, p.comm.rank);
    return r;
}
template <bool RP>
__device__ __forceinline__ void mv_store(const MatvecParams& p, const RpState& rp, int64_t idx, double v) {
    if constexpr (!RP) {
        store_out_d(p.y, p.y_dtype, idx, v);
    } else if (p.rp_mode == RP_ALLREDUCE) {
        for (int r = 0; r < p.comm.world; r++) reinterpret_cast<double*>(p.comm.peers[r] + rp.off)[idx] = v;
    } else {
        const float f = (float)v;
        for (int r = 0; r < p.comm.world; r++) reinterpret_cast<float*>(p.comm.peers[r] + rp.off)[idx] = f;
    }
}
// row-parallel shards each add their partial sum: the bias must enter the total once (rank 0)
template <bool RP>
__device__ __forceinline__ bool mv_use_bias(const MatvecParams& p) { return p.bias && (!RP || p.rp_mode != RP_ALLREDUCE || p.comm.rank == 0); }
constexpr int MV_BAR_DONE = 5;  // consumers + fix-up warp: every output store of this CTA is issued

// RP (compile time, like GRP: the plain matvec carries none of it): fused tensor-parallel exchange, producer side
template <class F, int MB, bool PRO, bool GRP, bool RP>
__global__ void __launch_bounds__(MV_THREADS, 1) matvec_kernel(const MatvecParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + MV_MAX_STAGES;
    uint8_t* xhat = smem + MV_HDR_BYTES;                  // fused prologue: quantised activation records [kc][m][320]
    uint8_t* stages = smem + MV_HDR_BYTES + (PRO ? p.xhat_bytes : 0);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t G = gridDim.x, g = blockIdx.x;
    if (p.trace && tid == 0) {
        p.trace[g * 8 + 0] = globaltimer_ns();
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[g * 8 + 2] = smid;
    }
    const int64_t c0 = sk_begin(g, p.C, G), c1 = sk_begin(g + 1, p.C, G);
    const int n_chunks = (int)(c1 - c0);
    const int KC = (int)p.KC;
    const int nst = p.nstages;
    SkPlan& sp = *reinterpret_cast<SkPlan*>(smem + 2 * MV_MAX_STAGES * 8);       // shared plan (computed once)

    if (tid == 0) {
        sp = sk_plan(c0, c1, p.KC);
        for (int s = 0; s < nst; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MV_CONSUMER_WARPS);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    pdl_launch_dependents();  // let the next kernel of the stream start its own weight prefetch
    __syncthreads();

    if (warp == MV_CONSUMER_WARPS) {
        // ===================== producer: one thread drives the TMA engine =====================
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            const uint32_t xbytes = PRO ? 0u : (uint32_t)p.M * ACT_REC_BYTES;
            const uint32_t wbytes = (uint32_t)p.chunk_bytes;
            constexpr bool grouped = GRP;  // expert-bank launch (compile-time: the plain matvec carries none of it)
            const int64_t cpw = (int64_t)p.tpw * p.KC;  // grouped: chunks per weight
            if (grouped) pdl_wait();                      // the expert selection is produced by the preceding kernel
            // j-th processed chunk -> (weight address, activation record address)
            auto chunk_vc = [&](int j) -> int64_t {      // chunk index in the (concatenated) chunk array
                if (j < sp.nH) return c0 + j;
                if (j < sp.nH + sp.nT) return c0 + sp.nH + sp.nF + (j - sp.nH);
                return c0 + sp.nH + (j - sp.nH - sp.nT);
            };
            auto chunk_src = [&](int j) -> const uint8_t* {
                const int64_t vc = chunk_vc(j);
                if (!grouped) return p.w + vc * (int64_t)wbytes;
                const int64_t slot = vc / cpw;
                const int e = p.sel[slot];
                return e < 0 ? nullptr : p.w_table[e] + (vc - slot * cpw) * (int64_t)wbytes;
            };
            auto chunk_x = [&](int j) -> const uint8_t* {
                int kc;
                if (j < sp.nH) kc = sp.kcH + j;
                else if (j < sp.nH + sp.nT) kc = j - sp.nH;
                else kc = (j - sp.nH - sp.nT) % KC;
                if (!grouped) return p.xq + (size_t)kc * xbytes;
                const int64_t slot = chunk_vc(j) / cpw;
                return p.xq + ((size_t)kc * p.x_rows + (size_t)(slot / p.x_slot_div)) * ACT_REC_BYTES;
            };
            // Phase 1 (before griddepcontrol.wait): weights do not depend on the preceding kernels.  Fill the
            // shared-memory ring and ask the TMA engine to pull the rest of this CTA's range into L2, so HBM
            // keeps streaming across the kernel boundary while the predecessor drains.
            const int pre = n_chunks < nst ? n_chunks : nst;
            // grouped mode: a slot whose selection is negative (expert not hosted by this rank) is skipped -- no weight
            // copy, a flag behind the stage's activation record tells the consumers to leave the accumulators at zero
            auto issue_w = [&](int j, uint8_t* st, uint64_t* bar) {
                const uint8_t* src = chunk_src(j);
                if (grouped) {
                    const bool skip = src == nullptr;
                    *reinterpret_cast<volatile int*>(st + wbytes + xbytes) = skip ? 1 : 0;
                    mbar_arrive_expect_tx(bar, (skip ? 0u : wbytes) + xbytes);
                    if (skip) return;
                } else {
                    mbar_arrive_expect_tx(bar, wbytes + xbytes);
                }
                bulk_g2s_hint(st, src, wbytes, bar, pol);
            };
            for (int j = 0; j < pre; j++) issue_w(j, stages + (size_t)j * p.stage_bytes, &full[j]);
            int npf = n_chunks - pre;
            if (npf > p.l2_prefetch_chunks) npf = p.l2_prefetch_chunks;
            for (int j = pre; j < pre + npf; j++)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(chunk_src(j)), "r"(wbytes) : "memory");
            pdl_wait();  // activations (written by the preceding kernel) are visible from here on
            if (p.trace) p.trace[g * 8 + 4] = globaltimer_ns();
            if (!PRO)
                for (int j = 0; j < pre; j++)
                    bulk_g2s(stages + (size_t)j * p.stage_bytes + wbytes, chunk_x(j), xbytes, &full[j]);
            int s = 0;
            uint32_t ph = 0;  // second use of each stage waits for the consumers' first release (phase 0)
            for (int j = pre; j < n_chunks; j++) {
                mbar_wait(&empty[s], ph);
                uint8_t* st = stages + (size_t)s * p.stage_bytes;
                issue_w(j, st, &full[s]);
                if (!PRO) bulk_g2s(st + wbytes, chunk_x(j), xbytes, &full[s]);
                if (++s == nst) { s = 0; ph ^= 1u; }
            }
            // every chunk of this launch is requested: keep HBM busy with the successor's first chunks (L2 prefetch)
            if (!GRP && p.next_pf > 0) {
                for (int64_t ng = g; ng < p.next_G; ng += G) {
                    const int64_t a0 = sk_begin(ng, p.next_C, p.next_G), a1 = sk_begin(ng + 1, p.next_C, p.next_G);
                    const SkPlan np_ = sk_plan(a0, a1, p.next_KC);
                    const int n = (int)(a1 - a0) < p.next_pf ? (int)(a1 - a0) : p.next_pf;
                    for (int j = 0; j < n; j++)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.next_w + sk_chunk_at(np_, a0, j) * (int64_t)p.next_chunk_bytes),
                                     "r"(p.next_chunk_bytes)
                                     : "memory");
                }
            }
        }
        return;
    }
    if (warp == MV_CONSUMER_WARPS + 1) {
        // ===================== fix-up warp: arrival atomics + ordered reduction of split tiles =====================
        // Runs beside the consumers (they only bar.arrive), so neither the math nor the TMA stream ever waits
        // for an atomic round trip.  Split tiles are processed first, so this finishes long before the CTA does.
        constexpr bool rp_on = RP;
        if (sp.nH == 0 && sp.nT == 0) {
            if (rp_on) named_bar_sync(MV_BAR_DONE, MV_CONSUMER_WARPS * 32 + 32);
            return;
        }
        pdl_wait();
        RpState rp{0u, 0};
        if constexpr (RP) rp = rp_begin(p);
        // arrival bookkeeping is computed before the barriers: only the atomic round trip is on the critical path
        int64_t tqs[2] = {sp.tH, sp.tT};
        int gfs[2], ncs[2], sgfs[2];
#pragma unroll
        for (int seg = 0; seg < 2; seg++) {
            const int64_t gf = sk_owner(tqs[seg] * p.KC, p.C, G), gl = sk_owner((tqs[seg] + 1) * p.KC - 1, p.C, G);
            gfs[seg] = (int)gf;
            ncs[seg] = (int)(gl - gf + 1);
            sgfs[seg] = (sk_begin(gf, p.C, G) == tqs[seg] * p.KC) ? 0 : 1;  // later contributors start inside the tile: slot 0
        }
#pragma unroll
        for (int seg = 0; seg < 2; seg++) {
            if ((seg == 0 ? sp.nH : sp.nT) == 0) continue;
            const int64_t tq = tqs[seg];
            const int gf = gfs[seg], gl = gfs[seg] + ncs[seg] - 1, nc = ncs[seg], sgf = sgfs[seg];
            named_bar_sync(2 + seg, MV_CONSUMER_WARPS * 32 + 32);  // every consumer warp stored its share of this tile
            if (p.trace && lane == 0) p.trace[g * 8 + 5] = globaltimer_ns();
            unsigned int old = 0;
            if (lane == 0) asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.ws_cnt + tq) : "memory");
            old = __shfl_sync(0xffffffffu, old, 0);
            if (p.trace && lane == 0) p.trace[g * 8 + 6] = globaltimer_ns();
            if (old != (unsigned int)(nc - 1)) continue;  // a later arriver reduces this tile
            // last arriver: sum the partials in CTA order (deterministic).  Every load of a batch of NB contributors
            // x all passes is issued before the first add, so the reduction costs one L2 round trip per NB contributors.
            const int64_t tql = GRP ? tq % p.tpw : tq;                      // tile index inside its weight
            const int64_t ybase = GRP ? (tq / p.tpw) * p.y_slot_stride : 0;  // grouped: output of slot tq / tpw
            constexpr int PASSES = 2 * MB;                       // 64 doubles (one double2 per lane) per pass
            constexpr int NB = MB == 1 ? 8 : (MB == 2 ? 4 : 2);  // contributors in flight
            double2 sum[PASSES];
#pragma unroll
            for (int v = 0; v < PASSES; v++) sum[v] = make_double2(0.0, 0.0);
            for (int g0 = gf; g0 <= gl; g0 += NB) {
                double2 tb[NB][PASSES];
#pragma unroll
                for (int u = 0; u < NB; u++) {
                    const int gg = g0 + u;
                    const double* src = p.ws_part + ((size_t)gg * 2 + (gg == gf ? sgf : 0)) * (TILE_ROWS * MB) + lane * 2;
#pragma unroll
                    for (int v = 0; v < PASSES; v++)
                        tb[u][v] = (gg <= gl) ? __ldcg(reinterpret_cast<const double2*>(src + v * 64)) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < NB; u++)
#pragma unroll
                    for (int v = 0; v < PASSES; v++) { sum[v].x += tb[u][v].x; sum[v].y += tb[u][v].y; }
            }
#pragma unroll
            for (int v = 0; v < PASSES; v++) {
                const int idx = (v * 32 + lane) * 2;
                const double sv[2] = {sum[v].x, sum[v].y};
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int rr = (idx + e) / MB, m = (idx + e) % MB;
                    const int64_t n = tql * TILE_ROWS + rr;
                    if (n < p.N && m < p.M) mv_store<RP>(p, rp, ybase + (int64_t)m * p.ldy + n, sv[e] + (mv_use_bias<RP>(p) ? (double)p.bias[n] : 0.0));
                }
            }
            if (lane == 0) p.ws_cnt[tq] = 0u;
            if (p.trace && lane == 0) p.trace[g * 8 + 7] = globaltimer_ns();
        }
        if (rp_on) named_bar_sync(MV_BAR_DONE, MV_CONSUMER_WARPS * 32 + 32);
        return;
    }

    // ===================== consumers =====================
    pdl_wait();  // y, the workspace and the bias may still be in use by the preceding kernel before this point
    RpState rp{0u, 0};
    if constexpr (RP) rp = rp_begin(p);
    const int g4 = lane >> 3, i = lane & 7;
    const FmtMeta meta{p.gpc};
    // f64 accumulators: every term is an exact product of an f32 scale and an integer partial, so the sum is
    // independent of the summation order up to 1e-16 -> the result is bit-reproducible against the oracle for any
    // grid size / stream-K split (oracle orc_matmul_q8).
    double acc[MV_STEPS][MB];
#pragma unroll
    for (int s4 = 0; s4 < MV_STEPS; s4++)
#pragma unroll
        for (int m = 0; m < MB; m++) acc[s4][m] = 0.0;

    if (PRO) fused_prologue<MB>(p, xhat, tid, warp, lane, g == 0);

    // segment walk: 0 = head (partial, slot 0), 1 = tail (partial, slot 1), 2 = full tiles
    int kcur = sp.nH > 0 ? sp.kcH : 0;  // k-chunk index of the chunk being processed (fused prologue addressing)
    int seg = sp.nH > 0 ? 0 : (sp.nT > 0 ? 1 : 2);
    int seg_left = seg == 0 ? sp.nH : (seg == 1 ? sp.nT : KC);  // chunks until the next flush
    int t = seg == 0 ? sp.tH : (seg == 1 ? sp.tT : sp.tF);
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_chunks; j++) {
        mbar_wait(&full[s], ph);
        if (p.trace && tid == 0 && j == 0) p.trace[g * 8 + 1] = globaltimer_ns();
        const uint8_t* wc = stages + (size_t)s * p.stage_bytes;
        const uint8_t* xr = PRO ? xhat + (size_t)kcur * p.M * ACT_REC_BYTES : wc + p.chunk_bytes;
        if (PRO) { if (++kcur == KC) kcur = 0; }

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
        const bool skip_chunk = GRP && *reinterpret_cast<const volatile int*>(wc + p.chunk_bytes + p.M * ACT_REC_BYTES) != 0;
        if (!(p.debug_flags & 1) && !skip_chunk)
#pragma unroll
        for (int s4 = 0; s4 < MV_STEPS; s4++) {
            const int r = MV_ROWS_PER_WARP * warp + 4 * s4 + g4;
            Unit u;
            F::template load_unit<true, F::NIB>(wc, r, i, u, meta);
#pragma unroll
            for (int m = 0; m < MB; m++) {
                int sA = 0, sB = 0;
                sA = __dp4a((int)u.v[0], (int)xa[m].x, sA); sA = __dp4a((int)u.v[1], (int)xa[m].y, sA);
                sA = __dp4a((int)u.v[2], (int)xa[m].z, sA); sA = __dp4a((int)u.v[3], (int)xa[m].w, sA);
                if constexpr (F::NIB) {  // bytes hold 16 x q (unsigned): exact u8 x s8 dot, one arithmetic shift back
                    sB = dp4a_us(u.v[4], xb[m].x, sB); sB = dp4a_us(u.v[5], xb[m].y, sB);
                    sB = dp4a_us(u.v[6], xb[m].z, sB); sB = dp4a_us(u.v[7], xb[m].w, sB);
                    sB >>= 4;
                } else {
                    sB = __dp4a((int)u.v[4], (int)xb[m].x, sB); sB = __dp4a((int)u.v[5], (int)xb[m].y, sB);
                    sB = __dp4a((int)u.v[6], (int)xb[m].z, sB); sB = __dp4a((int)u.v[7], (int)xb[m].w, sB);
                }
                sA -= u.off[0] * bsA[m];
                sB -= u.off[1] * bsB[m];
                double a_ = acc[s4][m];
                if (F::SUB == 32) {  // one scale per 32 weights
                    a_ = fma((double)__fmul_rn(u.a[0], dx[m]), (double)(sA + sB), a_);
                    if (F::HAS_MIN) a_ = fma(-(double)__fmul_rn(u.b[0], dx[m]), (double)(bsA[m] + bsB[m]), a_);
                } else {             // two 16-wide sub-blocks with their own scales
                    a_ = fma((double)__fmul_rn(u.a[0], dx[m]), (double)sA, a_);
                    if (F::HAS_MIN) a_ = fma(-(double)__fmul_rn(u.b[0], dx[m]), (double)bsA[m], a_);
                    a_ = fma((double)__fmul_rn(u.a[1], dx[m]), (double)sB, a_);
                    if (F::HAS_MIN) a_ = fma(-(double)__fmul_rn(u.b[1], dx[m]), (double)bsB[m], a_);
                }
                acc[s4][m] = a_;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == nst) { s = 0; ph ^= 1u; }

        if (--seg_left > 0) continue;

        // ---- segment / tile boundary: reduce the 8 lanes of each row and flush ----
#pragma unroll
        for (int s4 = 0; s4 < MV_STEPS; s4++)
#pragma unroll
            for (int m = 0; m < MB; m++) {
                double v = acc[s4][m];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                acc[s4][m] = v;
            }
        if (seg == 2) {
            if (i == 0) {
#pragma unroll
                const int64_t tl = GRP ? t % p.tpw : t;
                const int64_t ybase = GRP ? (int64_t)(t / p.tpw) * p.y_slot_stride : 0;
                for (int s4 = 0; s4 < MV_STEPS; s4++) {
                    const int64_t n = tl * TILE_ROWS + MV_ROWS_PER_WARP * warp + 4 * s4 + g4;
                    if (n < p.N) {
                        const double bv = mv_use_bias<RP>(p) ? (double)p.bias[n] : 0.0;
#pragma unroll
                        for (int m = 0; m < MB; m++)
                            if (m < p.M) mv_store<RP>(p, rp, ybase + (int64_t)m * p.ldy + n, acc[s4][m] + bv);
                    }
                }
            }
        } else {
            // partial tile: publish my share, then (warp 0) announce the arrival with a release atomic
            double* part = p.ws_part + ((size_t)g * 2 + seg) * (TILE_ROWS * MB);
            if (i == 0) {
#pragma unroll
                for (int s4 = 0; s4 < MV_STEPS; s4++) {
                    const int rr = MV_ROWS_PER_WARP * warp + 4 * s4 + g4;
#pragma unroll
                    for (int m = 0; m < MB; m++) part[rr * MB + m] = acc[s4][m];
                }
                __threadfence_block();
            }
            __syncwarp();
            asm volatile("bar.arrive %0, %1;" ::"r"(2 + seg), "r"(MV_CONSUMER_WARPS * 32 + 32) : "memory");
        }
#pragma unroll
        for (int s4 = 0; s4 < MV_STEPS; s4++)
#pragma unroll
            for (int m = 0; m < MB; m++) acc[s4][m] = 0.0;

        // ---- next segment / tile ----
        if (seg == 0 && sp.nT > 0) { seg = 1; seg_left = sp.nT; t = sp.tT; kcur = 0; }
        else if (seg != 2) { seg = 2; seg_left = KC; t = sp.tF; kcur = 0; }
        else { seg_left = KC; t++; }

    }
    if constexpr (RP) {
        // ---- fused exchange, producer side: this CTA's peer stores are all issued; the last CTA of the launch publishes ----
        named_bar_sync(MV_BAR_DONE, MV_CONSUMER_WARPS * 32 + 32);
        if (tid == 0) {
            uint8_t* mine = p.comm.peers[p.comm.rank];
            const bool ar = p.rp_mode == RP_ALLREDUCE;
            unsigned int* done = reinterpret_cast<unsigned int*>(mine + (ar ? COMM_OFF_AR_DONE : COMM_OFF_AG_DONE));
            __threadfence_system();  // the CTA's peer stores (ordered before this thread by the barrier) are performed system-wide
            unsigned int old;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(done) : "memory");
            if (old == (unsigned int)(G - 1)) {
                *done = 0u;  // next launch (ordered after this one by the stream)
                const int par = ar ? (int)(rp.epoch & 1u) : 0;
                const int foff = ar ? COMM_OFF_AR_FLAGS : COMM_OFF_AG_FLAGS;
                for (int r = 0; r < p.comm.world; r++)
                    st_release_sys(reinterpret_cast<unsigned int*>(p.comm.peers[r] + foff) + par * COMM_MAX_WORLD + p.comm.rank, rp.epoch);
                *reinterpret_cast<unsigned int*>(mine + (ar ? COMM_OFF_AR_EPOCH : COMM_OFF_AG_EPOCH)) = rp.epoch;
            }
        }
    }
    if (p.trace && tid == 0) p.trace[g * 8 + 3] = globaltimer_ns();
}

template <class F, int MB, bool PRO, bool GRP, bool RP = false>
static cudaError_t launch_t(const MatvecParams& p, int grid, int smem, cudaStream_t st) {
    static bool configured[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(matvec_kernel<F, MB, PRO, GRP, RP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(MV_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: overlap with the predecessor's tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, matvec_kernel<F, MB, PRO, GRP, RP>, p);
    if (le != cudaSuccess) return le;
    count_launch();
    return cudaGetLastError();
}

template <class F>
static cudaError_t launch_f(const MatvecParams& p, int mb, int grid, int smem, cudaStream_t st) {
    if (p.w_table) return mb == 1 && !p.pro && !p.rp_mode ? launch_t<F, 1, false, true>(p, grid, smem, st) : cudaErrorInvalidValue;
    if (p.rp_mode != RP_NONE) {  // fused tensor-parallel exchange (row-parallel o / down, vocabulary-parallel lm_head)
        if (p.pro) return cudaErrorInvalidValue;
        switch (mb) {
            case 1: return launch_t<F, 1, false, false, true>(p, grid, smem, st);
            case 2: return launch_t<F, 2, false, false, true>(p, grid, smem, st);
            case 4: return launch_t<F, 4, false, false, true>(p, grid, smem, st);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (mb) {
        case 1: return p.pro ? launch_t<F, 1, true, false>(p, grid, smem, st) : launch_t<F, 1, false, false>(p, grid, smem, st);
        case 2: return p.pro ? launch_t<F, 2, true, false>(p, grid, smem, st) : launch_t<F, 2, false, false>(p, grid, smem, st);
        case 4: return p.pro ? launch_t<F, 4, true, false>(p, grid, smem, st) : launch_t<F, 4, false, false>(p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}


}  // namespace b200q
