// gemm_tc.cu -- host side of the tcgen05 dequant-GEMM: activation staging, split-K reduction, tile plan, dispatch.
// The kernel lives in gemm_impl.cuh and is instantiated per format in inst_<format>.cu.
#include <cstdlib>

#include "gemm_impl.cuh"

namespace b200q {

// ------------------------------------------------------------------------------------------------
// activation staging: x[M,K] (f32/f16/bf16) -> f16 UMMA tiles  xs[mt][ks][Mt rows x 128 B, SW128]
// ------------------------------------------------------------------------------------------------
__global__ void stage_x_kernel(const void* __restrict__ x, int x_dtype, int64_t M, int64_t K, int64_t ldx, const int32_t* __restrict__ perm,
                               int Mt, int KS, uint8_t* __restrict__ xs) {
    pdl_launch_dependents();
    pdl_wait();  // x is written by the operator just ahead in the stream
    const int ks = blockIdx.x;
    const int64_t m = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;  // row in padded space
    const int kk = threadIdx.x;                                          // 0..63
    const int64_t mt = m / Mt;
    const int mr = (int)(m % Mt);
    // position kk of the sub-stage holds source element (kk & ~3) | {0,2,1,3}[kk & 3]  (see unit_to_f16)
    const int64_t k = (int64_t)ks * 64 + ((kk & ~3) | (((kk & 1) << 1) | ((kk >> 1) & 1)));
    float v = 0.0f;
    if (m < M && k < K) v = load_in(x, x_dtype, m * ldx + (perm ? perm[k] : k));
    v = fminf(fmaxf(v, -65504.0f), 65504.0f);
    uint8_t* dst = xs + ((size_t)(mt * KS + ks) * Mt) * 128 + (size_t)mr * 128 + ((((kk >> 3) ^ (mr & 7)) << 4) + ((kk & 7) << 1));
    *reinterpret_cast<__half*>(dst) = __float2half_rn(v);
}

// split-K reduction: y[m,n] = sum_s partial[s][m][n] (+ bias), summed in split order (deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t M, int64_t N, const float* __restrict__ bias, void* y,
                                     int y_dtype, int64_t ldy) {
    pdl_launch_dependents();
    pdl_wait();  // the partial sums of the GEMM just ahead
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const int64_t m = idx / N, n = idx % N;
    float acc = 0.0f;
    for (int s = 0; s < splits; s++) acc += partial[((size_t)s * M + m) * N + n];
    store_out(y, y_dtype, m * ldy + n, acc + (bias ? bias[n] : 0.0f));
}

// tail split-K reduction: y[m, n] of the tail tiles = sum over splits of partial_tail[tile][s][r][c] (+ bias), split order (deterministic)
__global__ void splitk_tail_reduce_kernel(const float* __restrict__ partial, int tail_first, int tail_n, int S, int MT, int Mt, int64_t M, int64_t N,
                                          const float* __restrict__ bias, void* y, int y_dtype, int64_t ldy) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (tile, c, r) with r fastest: coalesced along n
    const int64_t per = (int64_t)TILE_ROWS * Mt;
    if (idx >= (int64_t)tail_n * per) return;
    const int tl = (int)(idx / per), c = (int)((idx % per) / TILE_ROWS), r = (int)(idx % TILE_ROWS);
    const int lin = tail_first + tl, mt = lin % MT, t = lin / MT;
    const int64_t m = (int64_t)mt * Mt + c, n = (int64_t)t * TILE_ROWS + r;
    if (m >= M || n >= N) return;
    float acc = 0.0f;
    for (int s_ = 0; s_ < S; s_++) acc += partial[(((size_t)tl * S + s_) * TILE_ROWS + r) * Mt + c];
    store_out(y, y_dtype, m * ldy + n, acc + (bias ? bias[n] : 0.0f));
}


// every launch of this file carries the programmatic-dependent-launch attribute: the staging / GEMM / reduction kernels of a
// batched-decode step sit in the same PDL chain as the glue operators (a plain launch would serialise on both of its edges)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_g(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

// tail plan of a wide-M launch: (first tail tile, tail tiles, splits); splits == 1 means "no tail split"
static void tail_plan(const b200q_weight* w, int MT, int64_t M, int grid, int* first, int* n, int* S) {
    *first = 0; *n = 0; *S = 1;
    // Measured in round 2 (B200Q_GEMM_TAIL=1): correct, and 3-6 % SLOWER (14336x4096xM2048: 814 vs 864 TFLOP/s; 28672: 943 vs 992)
    // although it removes a 7th wave that is 95 % empty -- the 8 CTAs of that wave run alone at a much higher rate, i.e. the kernel
    // is bounded by a chip-wide resource (power / L2), not per-SM time, and the extra registers + partial traffic cost more than
    // the idle SMs.  Kept as an opt-in experiment.
    static const bool on = [] { const char* e_ = getenv("B200Q_GEMM_TAIL"); return e_ && atoi(e_) != 0; }();
    const int64_t tiles = w->T * MT;
    if (!on || M <= 128 || tiles <= grid || tiles % grid == 0) return;
    const int rem = (int)(tiles % grid);
    int s = grid / rem;                       // rem * s <= grid: the tail is ONE short wave
    if (s > (int)w->KC / 2) s = (int)w->KC / 2;  // >= 2 chunks per split
    if (s < 2) return;
    *first = (int)(tiles - rem); *n = rem; *S = s;
}

// split-K factor: the work items (tiles x splits) should fill whole waves of the SMs.  Wide-M launches only split
// when there are fewer tiles than SMs; skinny launches (one M tile, HBM-bound) pick the split with the least wave
// quantisation loss, since a 112-tile weight on 148 SMs otherwise idles a quarter of the machine.
static int pick_splits(const b200q_weight* w, int MT, int64_t M) {
    const int64_t tiles = w->T * MT;
    const int sms = w->num_sms;
    int smax = 8;
    if (smax > (int)w->KC / 2) smax = (int)w->KC / 2;
    if (smax < 1) smax = 1;
    // tuning knob (skinny launches only): B200Q_GEMM_SPLITS forces the split-K factor (clamped to [1, KC / 2])
    static const int forced = [] { const char* e_ = getenv("B200Q_GEMM_SPLITS"); return e_ ? atoi(e_) : 0; }();
    if (forced > 0 && M <= 128) return forced > smax ? smax : forced;
    if (M > 128) {
        int s = (int)(sms / tiles);
        if (s > smax) s = smax;
        return s < 1 ? 1 : s;
    }
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= smax; s++) {
        const int64_t items = tiles * s;
        const int64_t waves = (items + sms - 1) / sms;
        double eff = (double)items / (double)(waves * sms);
        // partial sums cost s x M x N f32 written and read again, relative to the packed weight bytes streamed once
        const double extra = s > 1 ? (double)s * (double)M * (double)w->N * 8.0 / (double)w->device_bytes : 0.0;
        eff = eff / (1.0 + extra) - 0.01 * (s - 1);
        if (eff > best_eff) { best_eff = eff; best = s; }
    }
    return best;
}

static int pick_mt(int64_t M, int* MT) {
    int64_t mtiles = (M + 255) / 256;
    int64_t per = (M + mtiles - 1) / mtiles;
    int Mt = (int)((per + 31) / 32 * 32);
    *MT = (int)mtiles;
    return Mt;
}

size_t gemm_ws_bytes(const b200q_weight* w, int64_t M) {
    if (M < 1) return 0;
    int MT;
    int Mt = pick_mt(M, &MT);
    size_t xs = ((size_t)MT * Mt * (size_t)w->K_pad * 2 + 255) & ~(size_t)255;
    int S = pick_splits(w, MT, M);
    int tf, tn, ts;
    tail_plan(w, MT, M, w->num_sms, &tf, &tn, &ts);
    return xs + (S > 1 ? (size_t)S * M * w->N * sizeof(float) : 0) + (ts > 1 ? (size_t)tn * ts * TILE_ROWS * Mt * sizeof(float) : 0);
}


static cudaError_t gemm_launch_family(int family, const GemmParams& p, int grid, int smem, cudaStream_t st) {
    switch (family) {
        case B200Q_FAM_Q4_K: return gemm_launch<B200Q_FAM_Q4_K>(p, grid, smem, st);
        case B200Q_FAM_Q6_K: return gemm_launch<B200Q_FAM_Q6_K>(p, grid, smem, st);
        case B200Q_FAM_Q8_0: return gemm_launch<B200Q_FAM_Q8_0>(p, grid, smem, st);
        case B200Q_FAM_Q5_K: return gemm_launch<B200Q_FAM_Q5_K>(p, grid, smem, st);
        case B200Q_FAM_Q4_1: return gemm_launch<B200Q_FAM_Q4_1>(p, grid, smem, st);
        case B200Q_FAM_Q5_1: return gemm_launch<B200Q_FAM_Q5_1>(p, grid, smem, st);
        case B200Q_FAM_Q2_K: return gemm_launch<B200Q_FAM_Q2_K>(p, grid, smem, st);
        case B200Q_FAM_Q3_K: return gemm_launch<B200Q_FAM_Q3_K>(p, grid, smem, st);
        case B200Q_FAM_IQ4_XS: return gemm_launch<B200Q_FAM_IQ4_XS>(p, grid, smem, st);
        case B200Q_FAM_TQ2_0: return gemm_launch<B200Q_FAM_TQ2_0>(p, grid, smem, st);
        case B200Q_FAM_I8S: return gemm_launch<B200Q_FAM_I8S>(p, grid, smem, st);
        case B200Q_FAM_G4: return gemm_launch<B200Q_FAM_G4>(p, grid, smem, st);
        default: return cudaErrorNotSupported;
    }
}

cudaError_t launch_gemm_tc(const b200q_weight* w, const void* x, int x_dtype, int64_t M, int64_t ldx, void* y, int y_dtype, int64_t ldy,
                           uint8_t* ws, size_t ws_bytes, cudaStream_t st) {
    (void)ws_bytes;
    int MT;
    const int Mt = pick_mt(M, &MT);
    const int KS = (int)w->KC * 4;
    // 1. stage activations as f16 UMMA tiles
    {
        dim3 block(64, 4);
        dim3 grid((unsigned)KS, (unsigned)((int64_t)MT * Mt / 4));
        cudaError_t e = launch_pdl_g(stage_x_kernel, grid, block, 0, st, x, x_dtype, M, w->K, ldx, w->perm, Mt, KS, ws);
        if (e != cudaSuccess) return e;
    }
    // 2. GEMM
    GemmParams p;
    p.w = w->data;
    p.xs = ws;
    p.y = y;
    p.bias = w->bias;
    p.N = w->N;
    p.M = M;
    p.ldy = ldy;
    p.y_dtype = y_dtype;
    p.T = (int)w->T;
    p.KC = (int)w->KC;
    p.MT = MT;
    p.Mt = Mt;
    p.gpc = w->gpc;
    p.chunk_bytes = w->chunk_bytes;
    p.w_stage_bytes = (w->chunk_bytes + 127) & ~127;
    p.xsub = Mt <= 64 ? 4 : 1;
    p.x_stage_bytes = Mt * 128 * p.xsub;
    // activation ring: deep enough to cover the L2 latency when the MMAs are short (small Mt), <= 128 KB
    int nx = (128 * 1024) / p.x_stage_bytes;
    if (nx > GT_NX) nx = GT_NX;
    if (nx < 4) nx = 4;
    if (p.xsub == 4) nx = Mt <= 32 ? 4 : 3;  // whole-chunk stages: keep the shared memory for the weight ring (the HBM stream)
    int nw_cap = GT_MAX_NW;
    // tuning knobs (defaults unchanged): B200Q_GEMM_NX / B200Q_GEMM_NW trade weight stages for activation stages
    if (const char* e = getenv("B200Q_GEMM_NX")) { int v = atoi(e); if (v >= 2 && v <= GT_NX) nx = v; }
    if (const char* e = getenv("B200Q_GEMM_NW")) { int v = atoi(e); if (v >= 2 && v <= GT_MAX_NW) nw_cap = v; }
    if (GT_HDR + nx * p.x_stage_bytes + 2 * p.w_stage_bytes > 227 * 1024) return cudaErrorNotSupported;
    p.nx = nx;
    int avail = 227 * 1024 - GT_HDR - nx * p.x_stage_bytes;
    int nw = avail / p.w_stage_bytes;
    if (nw > nw_cap) nw = nw_cap;
    if (nw < 2) return cudaErrorNotSupported;
    p.nw = nw;
    p.nslots = Mt > 64 ? 2 : 3;   // TMEM: 2 accumulators x Mt (Mt <= 128) or one of 256 columns, A slots of 128 columns on top
    p.a_col = 512 - 128 * p.nslots;
    p.idesc = (1u << 4) | ((uint32_t)(Mt >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32, K-major A/B, M=128, N=Mt
    int smem = GT_HDR + nx * p.x_stage_bytes + nw * p.w_stage_bytes;
    p.splits = pick_splits(w, MT, M);
    p.dq_warps = Mt <= 128 ? 16 : 8;
    p.partial = reinterpret_cast<float*>(ws + (((size_t)MT * Mt * (size_t)w->K_pad * 2 + 255) & ~(size_t)255));
    // wide M tiles (prefill): pairs of CTAs on adjacent weight tiles can share every activation stage by TMA multicast
    // (B200Q_GEMM_XMC=1).  Measured in round 2: correct (all GEMM parity tests, odd tile counts included) and NEUTRAL -- 862 vs 863
    // TFLOP/s on 14336x4096xM2048, 986 vs 979 on 28672x4096 -- so the activation tile's L2 -> SM traffic is not what bounds the
    // kernel (round 1's hypothesis); the 14336-row shape loses 14 % to wave quantisation (896 tiles on 148 SMs = 6.05 waves).  Off by default.
    static const bool xmc_on = [] { const char* e_ = getenv("B200Q_GEMM_XMC"); return e_ && atoi(e_) != 0; }();
    p.xmc = (xmc_on && Mt > 128 && p.xsub == 1 && w->T >= 2) ? 1 : 0;
    int64_t tiles = (int64_t)(p.xmc ? (p.T + 1) / 2 * 2 : p.T) * MT * p.splits;
    int grid = (int)(tiles < w->num_sms ? tiles : w->num_sms);
    if (p.xmc) grid &= ~1;
    p.tail_first = p.tail_n = 0;
    p.tail_splits = 1;
    p.partial_tail = nullptr;
    if (!p.xmc && p.splits == 1 && grid == w->num_sms) {
        tail_plan(w, MT, M, grid, &p.tail_first, &p.tail_n, &p.tail_splits);
        if (p.tail_splits > 1) p.partial_tail = p.partial;   // (splits == 1: the uniform split-K buffer is empty, the tail buffer starts there)
    }
    cudaError_t ge;
    ge = gemm_launch_family(w->family, p, grid, smem, st);
    if (ge == cudaSuccess && p.tail_splits > 1) {
        const int64_t total_t = (int64_t)p.tail_n * TILE_ROWS * Mt;
        return launch_pdl_g(splitk_tail_reduce_kernel, dim3((unsigned)((total_t + 255) / 256)), dim3(256), 0, st, p.partial_tail, p.tail_first, p.tail_n,
                            p.tail_splits, MT, Mt, M, w->N, w->bias, y, y_dtype, ldy);
    }
    if (ge != cudaSuccess || p.splits == 1) return ge;
    const int64_t total = M * w->N;
    return launch_pdl_g(splitk_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, p.partial, p.splits, M, w->N, w->bias, y, y_dtype, ldy);
}

}  // namespace b200q
