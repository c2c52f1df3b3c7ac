// dstep_impl.cuh -- EXPERIMENTAL (opt-in, B200Q_DSTEP=1; not yet run on hardware): a persistent "op list" kernel for
// the decode step.  One CTA per SM walks a program of ops -- add+RMSNorm+quantise, quantised matvec, SwiGLU+quantise --
// separated by grid barriers, while ONE producer thread per CTA keeps the TMA weight stream running across the
// barriers (weights depend on nothing), so the 3-5 us of fixed cost that every separate launch / kernel boundary
// pays today is replaced by a ~1 us barrier hidden behind a full shared-memory ring.
//
// The arithmetic is copied from the verified kernels (matvec_impl.cuh consumer loop + fix-up, decode_ops.cu norm /
// SwiGLU / quantiser): same operation order => same bits.  Differences: activation records of a matvec are pulled from
// global memory into shared memory once per op (all k-chunks), stages hold weights only, and every value produced by
// another CTA inside this launch is read with ld.global.cg (L2) after the barrier.
#pragma once
#include "matvec_common.cuh"

namespace b200q {

enum { DS_END = 0, DS_NORMQ = 1, DS_MATVEC = 2, DS_SWIGLUQ = 3, DS_ATTN = 4, DS_ARGMAX_A = 5, DS_ARGMAX_B = 6, DS_EMBED = 7 };

struct StepOp {
    int type, family, gpc, chunk_bytes;
    int M, y_dtype, KC, H, F, pad0;
    int64_t N, C, ldy;
    const uint8_t* w;
    const uint8_t* xq;        // MATVEC input records [KC][M][320]
    void* y;
    const float* bias;
    double* ws_part;
    unsigned int* ws_cnt;
    const float* h_in;        // NORMQ
    const float* delta;
    float* h_out;
    const float* norm_w;
    const float* gate_up;     // SWIGLUQ
    uint8_t* xq_out;          // NORMQ / SWIGLUQ / ATTN output records
    float eps;
    int pad1;
    // ATTN: qkv [M, (nh + 2 nkv) hd], cache [M][max_ctx][nkv][hd], rope [max_ctx][hd/2][2], pos [M]
    const float* qkv;
    const int* pos;
    float* cache_k;
    float* cache_v;
    const float* rope;
    int nh, nkv, hd, max_ctx;
    // ARGMAX_A / ARGMAX_B: logits [M, V] -> candidates [M][G] -> ids [M], pos[m]++ ; EMBED: h_out[m, :] = table[ids[m], :]
    const float* logits;
    float* cand_val;
    int* cand_idx;
    int64_t* ids;
    int* pos_inc;
    const __half* table;
    int V, pad2;
};

struct StepParams {
    const StepOp* ops;
    int n_ops, nstages, stage_bytes, xhat_bytes;
    unsigned int* bar_count;  // grid barrier: arrivals of the current generation
    unsigned int* bar_gen;    // grid barrier: generation
};

constexpr int DS_CONSUMERS = MV_CONSUMER_WARPS * 32;  // 512

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// sense-reversing grid barrier, called by ONE thread per CTA; self-resetting, reusable across launches
__device__ __forceinline__ void grid_barrier(unsigned int* count, unsigned int* gen, unsigned int& my_gen, unsigned int G) {
    __threadfence();
    const unsigned int old = atomicAdd(count, 1u);
    if (old == G - 1) {
        atomicExch(count, 0u);
        __threadfence();
        atomicAdd(gen, 1u);
    } else {
        while (ld_acquire_gpu(gen) == my_gen) {
        }
    }
    my_gen++;
    __threadfence();
}

// quantise 256 values held one per thread (t < 256) into a record -- decode_ops.cu quant_store_record
__device__ __forceinline__ void ds_quant_store_record(float v, uint8_t* rec, int t) {
    const int wid = t >> 5, lane = t & 31;
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
    rec[t] = (uint8_t)(int8_t)q;
    if (lane == 0) {
        reinterpret_cast<float*>(rec + 256)[wid] = d;
        reinterpret_cast<uint32_t*>(rec + 288)[wid] = ((uint32_t)s & 0xFFFFu) | ((uint32_t)s_hi << 16);
    }
}

// ---- consumer side of one MATVEC op (512 threads); stage / phase state of the weight ring is carried across ops ----
template <class F, int MB>
__device__ __forceinline__ void ds_consume(const StepOp& op, const StepParams& p, const SkPlan& sp, int n_chunks, int64_t g, uint8_t* stages,
                                           const uint8_t* xhat, uint64_t* full, uint64_t* empty, int& s, uint32_t& ph, int tid, int warp, int lane) {
    const int g4 = lane >> 3, i = lane & 7;
    const FmtMeta meta{op.gpc};
    const int KC = op.KC, nst = p.nstages;
    double acc[MV_STEPS][MB];
#pragma unroll
    for (int s4 = 0; s4 < MV_STEPS; s4++)
#pragma unroll
        for (int m = 0; m < MB; m++) acc[s4][m] = 0.0;
    int kcur = sp.nH > 0 ? sp.kcH : 0;
    int seg = sp.nH > 0 ? 0 : (sp.nT > 0 ? 1 : 2);
    int seg_left = seg == 0 ? sp.nH : (seg == 1 ? sp.nT : KC);
    int t = seg == 0 ? sp.tH : (seg == 1 ? sp.tT : sp.tF);
    for (int j = 0; j < n_chunks; j++) {
        mbar_wait(&full[s], ph);
        const uint8_t* wc = stages + (size_t)s * p.stage_bytes;
        const uint8_t* xr = xhat + (size_t)kcur * op.M * ACT_REC_BYTES;
        if (++kcur == KC) kcur = 0;
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
        for (int s4 = 0; s4 < MV_STEPS; s4++) {
            const int r = MV_ROWS_PER_WARP * warp + 4 * s4 + g4;
            Unit u;
            F::template load_unit<true, F::NIB>(wc, r, i, u, meta);
#pragma unroll
            for (int m = 0; m < MB; m++) {
                int sA = 0, sB = 0;
                sA = __dp4a((int)u.v[0], (int)xa[m].x, sA); sA = __dp4a((int)u.v[1], (int)xa[m].y, sA);
                sA = __dp4a((int)u.v[2], (int)xa[m].z, sA); sA = __dp4a((int)u.v[3], (int)xa[m].w, sA);
                if constexpr (F::NIB) {
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
                if (F::SUB == 32) {
                    a_ = fma((double)__fmul_rn(u.a[0], dx[m]), (double)(sA + sB), a_);
                    if (F::HAS_MIN) a_ = fma(-(double)__fmul_rn(u.b[0], dx[m]), (double)(bsA[m] + bsB[m]), a_);
                } else {
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
                for (int s4 = 0; s4 < MV_STEPS; s4++) {
                    const int64_t n = (int64_t)t * TILE_ROWS + MV_ROWS_PER_WARP * warp + 4 * s4 + g4;
                    if (n < op.N) {
                        const double bv = op.bias ? (double)op.bias[n] : 0.0;
#pragma unroll
                        for (int m = 0; m < MB; m++)
                            if (m < op.M) store_out_d(op.y, op.y_dtype, (int64_t)m * op.ldy + n, acc[s4][m] + bv);
                    }
                }
            }
        } else {
            double* part = op.ws_part + ((size_t)g * 2 + seg) * (TILE_ROWS * MB);
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
            asm volatile("bar.arrive %0, %1;" ::"r"(2 + seg), "r"(DS_CONSUMERS + 32) : "memory");
        }
#pragma unroll
        for (int s4 = 0; s4 < MV_STEPS; s4++)
#pragma unroll
            for (int m = 0; m < MB; m++) acc[s4][m] = 0.0;
        if (seg == 0 && sp.nT > 0) { seg = 1; seg_left = sp.nT; t = sp.tT; kcur = 0; }
        else if (seg != 2) { seg = 2; seg_left = KC; t = sp.tF; kcur = 0; }
        else { seg_left = KC; t++; }
    }
    (void)tid;
}

// ---- fix-up warp of one MATVEC op: arrival atomics + ordered reduction of split tiles (matvec_impl.cuh) ----
template <int MB>
__device__ __forceinline__ void ds_fixup(const StepOp& op, const SkPlan& sp, int64_t g, int64_t G, int lane) {
    if (sp.nH == 0 && sp.nT == 0) return;
    int64_t tqs[2] = {sp.tH, sp.tT};
    int gfs[2], ncs[2], sgfs[2];
#pragma unroll
    for (int seg = 0; seg < 2; seg++) {
        const int64_t gf = sk_owner(tqs[seg] * op.KC, op.C, G), gl = sk_owner((tqs[seg] + 1) * op.KC - 1, op.C, G);
        gfs[seg] = (int)gf;
        ncs[seg] = (int)(gl - gf + 1);
        sgfs[seg] = (sk_begin(gf, op.C, G) == tqs[seg] * op.KC) ? 0 : 1;
    }
#pragma unroll
    for (int seg = 0; seg < 2; seg++) {
        if ((seg == 0 ? sp.nH : sp.nT) == 0) continue;
        const int64_t tq = tqs[seg];
        const int gf = gfs[seg], gl = gfs[seg] + ncs[seg] - 1, nc = ncs[seg], sgf = sgfs[seg];
        named_bar_sync(2 + seg, DS_CONSUMERS + 32);
        unsigned int old = 0;
        if (lane == 0) asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(op.ws_cnt + tq) : "memory");
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old != (unsigned int)(nc - 1)) continue;
        constexpr int PASSES = 2 * MB;
        constexpr int NB = MB == 1 ? 8 : (MB == 2 ? 4 : 2);
        double2 sum[PASSES];
#pragma unroll
        for (int v = 0; v < PASSES; v++) sum[v] = make_double2(0.0, 0.0);
        for (int g0 = gf; g0 <= gl; g0 += NB) {
            double2 tb[NB][PASSES];
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const int gg = g0 + u;
                const double* src = op.ws_part + ((size_t)gg * 2 + (gg == gf ? sgf : 0)) * (TILE_ROWS * MB) + lane * 2;
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
                const int64_t n = tq * TILE_ROWS + rr;
                if (n < op.N && m < op.M) store_out_d(op.y, op.y_dtype, (int64_t)m * op.ldy + n, sv[e] + (op.bias ? (double)op.bias[n] : 0.0));
            }
        }
        if (lane == 0) op.ws_cnt[tq] = 0u;
    }
}

// ---- ATTN op: port of decode_ops.cu attn_decode_kernel<HD> to the 512 consumer threads of one CTA (named barrier 4,
// scratch carved from the record region, qkv read through L2).  Same operation order => same bits. ----
template <int HD>
__device__ __forceinline__ void ds_attention(const StepOp& op, int head, int m, uint8_t* scratch, int t) {
    const int warp = t >> 5, lane = t & 31;
    constexpr int NT = DS_CONSUMERS, NW = NT / 32;
    constexpr int QE = HD / 4, EG = HD / 4, JG = NT / EG;
    double* s_o = reinterpret_cast<double*>(scratch);                  // [JG][HD]
    double* s_redd = s_o + JG * HD;                                      // [NW]
    float* s_redf = reinterpret_cast<float*>(s_redd + NW);               // [NW]
    float* sq = s_redf + NW;                                             // [HD] x 3, 16-byte aligned (NW = 16)
    float* sk = sq + HD;
    float* sv = sk + HD;
    float* s_sc = sv + HD;                                               // [max_ctx]
    const int nh = op.nh, nkv = op.nkv, max_ctx = op.max_ctx, M = op.M;
    const int kvh = head / (nh / nkv);
    const int p = op.pos[m];
    const int row = (nh + 2 * nkv) * HD;
    const float* qsrc = op.qkv + (size_t)m * row + (size_t)head * HD;
    const float* ksrc = op.qkv + (size_t)m * row + (size_t)(nh + kvh) * HD;
    const float* vsrc = op.qkv + (size_t)m * row + (size_t)(nh + nkv + kvh) * HD;
    float* ck = op.cache_k + ((size_t)m * max_ctx) * nkv * HD + (size_t)kvh * HD;
    float* cv = op.cache_v + ((size_t)m * max_ctx) * nkv * HD + (size_t)kvh * HD;
    const size_t pstride = (size_t)nkv * HD;
    const float* rt = op.rope + (size_t)p * HD;
    if (t < HD / 2) {
        const float c = rt[2 * t], sn = rt[2 * t + 1];
        const float q0 = __ldcg(qsrc + 2 * t), q1 = __ldcg(qsrc + 2 * t + 1);
        sq[2 * t] = __fsub_rn(__fmul_rn(q0, c), __fmul_rn(q1, sn));
        sq[2 * t + 1] = __fadd_rn(__fmul_rn(q0, sn), __fmul_rn(q1, c));
        const float k0 = __ldcg(ksrc + 2 * t), k1 = __ldcg(ksrc + 2 * t + 1);
        const float r0 = __fsub_rn(__fmul_rn(k0, c), __fmul_rn(k1, sn)), r1 = __fadd_rn(__fmul_rn(k0, sn), __fmul_rn(k1, c));
        sk[2 * t] = r0; sk[2 * t + 1] = r1;
        const float v0 = __ldcg(vsrc + 2 * t), v1 = __ldcg(vsrc + 2 * t + 1);
        sv[2 * t] = v0; sv[2 * t + 1] = v1;
        if (head % (nh / nkv) == 0) {
            ck[(size_t)p * pstride + 2 * t] = r0; ck[(size_t)p * pstride + 2 * t + 1] = r1;
            cv[(size_t)p * pstride + 2 * t] = v0; cv[(size_t)p * pstride + 2 * t + 1] = v1;
        }
    }
    named_bar_sync(4, NT);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn((float)HD));
    float lmax = -INFINITY;
    {
        const int part = t & 3;
        for (int j0 = 0; j0 <= p; j0 += NT / 4) {
            const int j = j0 + (t >> 2);
            double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
            if (j <= p) {
                const float* kr = (j < p) ? ck + (size_t)j * pstride + part * QE : sk + part * QE;
                const float* qq = sq + part * QE;
#pragma unroll
                for (int e = 0; e < QE; e += 4) {
                    const float4 kk = *reinterpret_cast<const float4*>(kr + e);
                    d0 = fma((double)qq[e + 0], (double)kk.x, d0); d1 = fma((double)qq[e + 1], (double)kk.y, d1);
                    d2 = fma((double)qq[e + 2], (double)kk.z, d2); d3 = fma((double)qq[e + 3], (double)kk.w, d3);
                }
            }
            double d = (d0 + d1) + (d2 + d3);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            if (j <= p) {
                const float sc = __fmul_rn((float)d, scale);
                if (part == 0) s_sc[j] = sc;
                lmax = fmaxf(lmax, sc);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
    if (lane == 0) s_redf[warp] = lmax;
    named_bar_sync(4, NT);
    float gmax = s_redf[0];
#pragma unroll
    for (int i = 1; i < NW; i++) gmax = fmaxf(gmax, s_redf[i]);
    double lsum = 0.0;
    for (int j = t; j <= p; j += NT) {
        const float pj = det_expf(__fsub_rn(s_sc[j], gmax));
        s_sc[j] = pj;
        lsum += (double)pj;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
    if (lane == 0) s_redd[warp] = lsum;
    named_bar_sync(4, NT);
    double dsum = 0.0;
#pragma unroll
    for (int i = 0; i < NW; i++) dsum += s_redd[i];
    const float den = (float)dsum;
    {
        const int eg = t % EG, jg = t / EG;
        double o0 = 0.0, o1 = 0.0, o2 = 0.0, o3 = 0.0;
        for (int j = jg; j <= p; j += JG) {
            const float pj = s_sc[j];
            const float4 vv = (j < p) ? *reinterpret_cast<const float4*>(cv + (size_t)j * pstride + 4 * eg) : *reinterpret_cast<const float4*>(sv + 4 * eg);
            o0 = fma((double)pj, (double)vv.x, o0); o1 = fma((double)pj, (double)vv.y, o1);
            o2 = fma((double)pj, (double)vv.z, o2); o3 = fma((double)pj, (double)vv.w, o3);
        }
        s_o[jg * HD + 4 * eg + 0] = o0; s_o[jg * HD + 4 * eg + 1] = o1; s_o[jg * HD + 4 * eg + 2] = o2; s_o[jg * HD + 4 * eg + 3] = o3;
        named_bar_sync(4, NT);
    }
    if (t < HD) {
        double o = 0.0;
#pragma unroll
        for (int i = 0; i < JG; i++) o += s_o[i * HD + t];
        const float outv = __fdiv_rn((float)o, den);
        const int kglob = head * HD + t;
        const int kc = kglob / CHUNK_K, tin = kglob % CHUNK_K;
        uint8_t* rec = op.xq_out + ((size_t)kc * M + m) * ACT_REC_BYTES;
        float amax = fabsf(outv);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
        const float d = __fdiv_rn(amax, 127.0f);
        const float id = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
        const int q = (int)roundf(__fmul_rn(outv, id));
        int s = q;
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const int s_hi = __shfl_sync(0xffffffffu, s, 16);
        rec[tin] = (uint8_t)(int8_t)q;
        if (lane == 0) {
            reinterpret_cast<float*>(rec + 256)[tin >> 5] = d;
            reinterpret_cast<uint32_t*>(rec + 288)[tin >> 5] = ((uint32_t)s & 0xFFFFu) | ((uint32_t)s_hi << 16);
        }
    }
    named_bar_sync(4, NT);   // the scratch is reused by the next (head, m) of this CTA
}

template <int MB>
__device__ __forceinline__ void ds_consume_family(const StepOp& op, const StepParams& p, const SkPlan& sp, int n_chunks, int64_t g, uint8_t* stages,
                                                  const uint8_t* xhat, uint64_t* full, uint64_t* empty, int& s, uint32_t& ph, int tid, int warp, int lane) {
    switch (op.family) {  // per op, outside the chunk loop: the loops themselves are fully specialised
        case B200Q_FAM_Q4_K: ds_consume<FmtQ4K, MB>(op, p, sp, n_chunks, g, stages, xhat, full, empty, s, ph, tid, warp, lane); break;
        case B200Q_FAM_Q6_K: ds_consume<FmtQ6K, MB>(op, p, sp, n_chunks, g, stages, xhat, full, empty, s, ph, tid, warp, lane); break;
        case B200Q_FAM_Q8_0: ds_consume<FmtQ8_0, MB>(op, p, sp, n_chunks, g, stages, xhat, full, empty, s, ph, tid, warp, lane); break;
        default: ds_consume<FmtG4, MB>(op, p, sp, n_chunks, g, stages, xhat, full, empty, s, ph, tid, warp, lane); break;
    }
}

// Shared memory: [barriers 256 B][red 16 doubles + inv][xhat (records of the current matvec, all k-chunks)][weight ring]
template <int MB>
__global__ void __launch_bounds__(MV_THREADS, 1) dstep_kernel(const StepParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + MV_MAX_STAGES;
    double* red = reinterpret_cast<double*>(smem + 256);         // [MB][16]
    float* s_inv = reinterpret_cast<float*>(smem + 256 + MB * 16 * 8);
    uint8_t* xhat = smem + 1024;
    uint8_t* stages = xhat + p.xhat_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t G = gridDim.x, g = blockIdx.x;
    const int nst = p.nstages;
    if (tid == 0) {
        for (int s = 0; s < nst; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MV_CONSUMER_WARPS);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    pdl_launch_dependents();
    __syncthreads();

    if (warp == MV_CONSUMER_WARPS) {
        // ===================== producer: streams the weights of every MATVEC op, in program order =====================
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            int s = 0, issued = 0;
            uint32_t ph = 0;
            for (int k = 0; k < p.n_ops; k++) {
                const StepOp& op = p.ops[k];
                if (op.type != DS_MATVEC) continue;
                int64_t Gk = G > op.C ? op.C : G;
                if (g >= Gk) continue;
                const int64_t c0 = sk_begin(g, op.C, Gk), c1 = sk_begin(g + 1, op.C, Gk);
                const SkPlan sp = sk_plan(c0, c1, op.KC);
                const int n = (int)(c1 - c0);
                const uint32_t wbytes = (uint32_t)op.chunk_bytes;
                for (int j = 0; j < n; j++) {
                    int64_t vc;
                    if (j < sp.nH) vc = c0 + j;
                    else if (j < sp.nH + sp.nT) vc = c0 + sp.nH + sp.nF + (j - sp.nH);
                    else vc = c0 + sp.nH + (j - sp.nH - sp.nT);
                    if (issued >= nst) mbar_wait(&empty[s], ph ^ 1u);   // the consumers' release of this stage's previous use
                    mbar_arrive_expect_tx(&full[s], wbytes);
                    bulk_g2s_hint(stages + (size_t)s * p.stage_bytes, op.w + vc * (int64_t)wbytes, wbytes, &full[s], pol);
                    issued++;
                    if (++s == nst) { s = 0; ph ^= 1u; }
                }
            }
        }
        return;
    }

    // consumers (warps 0..15) and the fix-up warp (17) walk the program together
    const bool is_fix = warp == MV_CONSUMER_WARPS + 1;
    unsigned int my_gen = 0;
    pdl_wait();   // inputs written by kernels ahead of this launch; an earlier launch of this program has passed its last barrier
    if (!is_fix && tid == 0) my_gen = ld_acquire_gpu(p.bar_gen);
    int s = 0;
    uint32_t ph = 0;
    for (int k = 0; k < p.n_ops; k++) {
        const StepOp& op = p.ops[k];
        if (op.type == DS_MATVEC) {
            int64_t Gk = G > op.C ? op.C : G;
            const bool active = g < Gk;
            SkPlan sp = {};
            int n_chunks = 0;
            if (active) {
                const int64_t c0 = sk_begin(g, op.C, Gk), c1 = sk_begin(g + 1, op.C, Gk);
                sp = sk_plan(c0, c1, op.KC);
                n_chunks = (int)(c1 - c0);
            }
            if (is_fix) {
                if (active) ds_fixup<MB>(op, sp, g, Gk, lane);
            } else {
                if (active) {
                    // activation records of every k-chunk -> shared memory (written by another CTA: L2 loads)
                    const int nvec = op.KC * op.M * (ACT_REC_BYTES / 16);
                    const uint4* src = reinterpret_cast<const uint4*>(op.xq);
                    uint4* dst = reinterpret_cast<uint4*>(xhat);
                    for (int v = tid; v < nvec; v += DS_CONSUMERS) dst[v] = __ldcg(src + v);
                    named_bar_sync(4, DS_CONSUMERS);
                    ds_consume_family<MB>(op, p, sp, n_chunks, g, stages, xhat, full, empty, s, ph, tid, warp, lane);
                }
            }
        } else if (!is_fix && op.type == DS_NORMQ) {
            // h_out = h_in (+ delta); xq_out = quant(rmsnorm(h_out) * w): CTA c handles k-chunk c (decode_ops.cu arithmetic)
            const int H = op.H;
            for (int kc = (int)g; kc < H / CHUNK_K; kc += (int)G) {
                for (int m = 0; m < op.M; m++) {
                    const float* hr = op.h_in + (size_t)m * H;
                    const float* dr = op.delta ? op.delta + (size_t)m * H : nullptr;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    for (int k4 = tid; k4 < H / 4; k4 += DS_CONSUMERS) {
                        float4 v = __ldcg(reinterpret_cast<const float4*>(hr) + k4);
                        if (dr) {
                            const float4 d4 = __ldcg(reinterpret_cast<const float4*>(dr) + k4);
                            v.x = __fadd_rn(v.x, d4.x); v.y = __fadd_rn(v.y, d4.y); v.z = __fadd_rn(v.z, d4.z); v.w = __fadd_rn(v.w, d4.w);
                        }
                        s0 = fma((double)v.x, (double)v.x, s0); s1 = fma((double)v.y, (double)v.y, s1);
                        s2 = fma((double)v.z, (double)v.z, s2); s3 = fma((double)v.w, (double)v.w, s3);
                    }
                    double ss = (s0 + s1) + (s2 + s3);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                    if (lane == 0) red[m * 16 + warp] = ss;
                    named_bar_sync(4, DS_CONSUMERS);
                    if (tid == 0) {
                        double tot = 0.0;
                        for (int w_ = 0; w_ < MV_CONSUMER_WARPS; w_++) tot += red[m * 16 + w_];
                        s_inv[m] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn((float)tot, (float)H), op.eps)));
                    }
                    named_bar_sync(4, DS_CONSUMERS);
                    if (tid < CHUNK_K) {
                        const int kk = kc * CHUNK_K + tid;
                        float mine = __ldcg(hr + kk);
                        if (dr) mine = __fadd_rn(mine, __ldcg(dr + kk));
                        op.h_out[(size_t)m * H + kk] = mine;
                        const float v = __fmul_rn(__fmul_rn(mine, s_inv[m]), op.norm_w[kk]);
                        ds_quant_store_record(v, op.xq_out + ((size_t)kc * op.M + m) * ACT_REC_BYTES, tid);
                    }
                    named_bar_sync(4, DS_CONSUMERS);
                }
            }
        } else if (!is_fix && op.type == DS_SWIGLUQ) {
            const int F = op.F;
            for (int kc = (int)g; kc < (F + CHUNK_K - 1) / CHUNK_K; kc += (int)G) {
                if (tid < CHUNK_K) {
                    for (int m = 0; m < op.M; m++) {
                        const int kk = kc * CHUNK_K + tid;
                        float v = 0.0f;
                        if (kk < F) {
                            const float gv = __ldcg(op.gate_up + (size_t)m * 2 * F + kk), uv = __ldcg(op.gate_up + (size_t)m * 2 * F + F + kk);
                            v = __fmul_rn(__fdiv_rn(gv, __fadd_rn(1.0f, det_expf(-gv))), uv);
                        }
                        ds_quant_store_record(v, op.xq_out + ((size_t)kc * op.M + m) * ACT_REC_BYTES, tid);
                    }
                }
            }
        }
        else if (!is_fix && op.type == DS_ATTN) {
            for (int hm = (int)g; hm < op.nh * op.M; hm += (int)G) {
                const int head = hm % op.nh, m = hm / op.nh;
                if (op.hd == 128) ds_attention<128>(op, head, m, xhat, tid);
                else ds_attention<64>(op, head, m, xhat, tid);
            }
        } else if (!is_fix && op.type == DS_ARGMAX_A) {
            // candidates of this CTA's slice of the vocabulary (lowest index wins ties, as argmax_kernel)
            const int V = op.V;
            const int lo = (int)((int64_t)g * V / G), hi = (int)((int64_t)(g + 1) * V / G);
            float* sv_ = reinterpret_cast<float*>(xhat);
            int* si_ = reinterpret_cast<int*>(xhat + 64);
            for (int m = 0; m < op.M; m++) {
                const float* rowp = op.logits + (size_t)m * V;
                float best = -INFINITY;
                int bi = 0x7fffffff;
                for (int i = lo + tid; i < hi; i += DS_CONSUMERS) {
                    const float v = __ldcg(rowp + i);
                    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                if (lane == 0) { sv_[warp] = best; si_[warp] = bi; }
                named_bar_sync(4, DS_CONSUMERS);
                if (tid == 0) {
                    for (int w_ = 1; w_ < MV_CONSUMER_WARPS; w_++)
                        if (sv_[w_] > best || (sv_[w_] == best && si_[w_] < bi)) { best = sv_[w_]; bi = si_[w_]; }
                    op.cand_val[(size_t)m * G + g] = best;
                    op.cand_idx[(size_t)m * G + g] = bi;
                }
                named_bar_sync(4, DS_CONSUMERS);
            }
        } else if (!is_fix && op.type == DS_ARGMAX_B) {
            if (g == 0 && tid < op.M) {
                const int m = tid;
                float best = -INFINITY;
                int bi = 0x7fffffff;
                for (int c = 0; c < (int)G; c++) {
                    const float v = __ldcg(op.cand_val + (size_t)m * G + c);
                    const int i = __ldcg(op.cand_idx + (size_t)m * G + c);
                    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
                }
                op.ids[m] = bi;
                if (op.pos_inc) op.pos_inc[m] += 1;
            }
        } else if (!is_fix && op.type == DS_EMBED) {
            const int H = op.H;
            for (int idx = (int)g * DS_CONSUMERS + tid; idx < op.M * H; idx += (int)G * DS_CONSUMERS) {
                const int m = idx / H, kk = idx % H;
                op.h_out[(size_t)m * H + kk] = __half2float(op.table[(size_t)op.ids[m] * H + kk]);
            }
        }
        // ---- end of op: every CTA's outputs (consumer stores + fix-up stores) are complete before anyone starts the next ----
        if (k + 1 < p.n_ops) {
            named_bar_sync(5, DS_CONSUMERS + 32);   // consumers + fix-up warp of this CTA are done with op k
            if (!is_fix) {
                if (tid == 0) grid_barrier(p.bar_count, p.bar_gen, my_gen, (unsigned int)G);
                named_bar_sync(6, DS_CONSUMERS);
            }
        }
    }
}

}  // namespace b200q
