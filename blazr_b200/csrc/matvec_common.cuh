// matvec_common.cuh -- parameters and stream-K helpers shared by the matvec kernel and its host dispatcher.
#pragma once
#include "formats.cuh"
#include "internal.h"

namespace b200q {


constexpr int MV_CONSUMER_WARPS = 16;
constexpr int MV_THREADS = (MV_CONSUMER_WARPS + 2) * 32;  // + producer warp + fix-up warp
constexpr int MV_ROWS_PER_WARP = TILE_ROWS / MV_CONSUMER_WARPS;  // 8
constexpr int MV_STEPS = MV_ROWS_PER_WARP / 4;                    // 2 (4 rows per step, 8 lanes per row)
constexpr int MV_MAX_STAGES = 10;
constexpr int MV_HDR_BYTES = 256;  // barriers + flags

struct MatvecParams {
    const uint8_t* w;
    const uint8_t* xq;
    void* y;
    const float* bias;
    double* ws_part;
    unsigned int* ws_cnt;
    int64_t N;
    int M, y_dtype;
    int64_t ldy;
    int64_t KC, C;
    int gpc, nstages, chunk_bytes, stage_bytes;
    long long* trace;  // debug: 4 x globaltimer per CTA (null in production)
    int debug_flags;   // debug: bit0 = consumers skip the math (measures the pure TMA stream)
    int l2_prefetch_chunks;  // per CTA: chunks beyond the smem ring to pull into L2 before griddepcontrol.wait
    // fused activation producers (prologue): 0 = records arrive by TMA from xq, 1 = quant(rmsnorm(h_in + delta) * nw),
    // 2 = quant(silu(gate) * up).  The whole quantised activation then lives in shared memory for the CTA's life.
    int pro;
    int xhat_bytes;
    const float* h_in;
    const float* delta;
    float* h_out;
    const float* norm_w;
    float eps;
    const float* gate_up;
    // grouped (expert bank) mode, w_table != null: the launch runs n_slots independent M = 1 matvecs; slot s uses the
    // weight w_table[sel[s]], the activation record row s / x_slot_div of xq (x_rows rows per k-chunk) and writes
    // y[s * y_slot_stride + n].  Tiles are numbered slot-major (tile t -> slot t / tpw), so the stream-K split, the
    // partial slots and the fix-up work unchanged on the concatenated chunk range.
    const uint8_t* const* w_table;
    const int32_t* sel;
    int tpw;              // tiles per weight
    int x_rows, x_slot_div;
    int64_t y_slot_stride;
};

__device__ __forceinline__ int64_t sk_begin(int64_t g, int64_t C, int64_t G) { return g * C / G; }
__device__ __forceinline__ int64_t sk_owner(int64_t c, int64_t C, int64_t G) { return ((c + 1) * G - 1) / C; }

// Processing order of a CTA's chunk range [c0,c1): the two tiles it shares with its neighbours first
// (head = tail end of tile t_first, then tail = first chunks of tile t_last), the tiles it owns entirely last.
// Both contributors of a split tile therefore finish their share early in their lifetime and the
// fix-up (atomic arrival + ordered reduction by the last arriver) happens mid-stream instead of in the tail.
struct SkPlan {
    int nH, nT, nF;          // chunks in the head / tail / full segments
    int kcH;                 // k-chunk index at which the head segment starts (tail and full start at 0)
    int tH, tT, tF;          // tile indices: head tile, tail tile, first full tile
};
__device__ __forceinline__ SkPlan sk_plan(int64_t c0, int64_t c1, int64_t KC) {
    SkPlan s;
    const int64_t t0 = c0 / KC, t1 = (c1 - 1) / KC;
    const int kc0 = (int)(c0 - t0 * KC);
    const int64_t head_end = (kc0 != 0 || c1 < (t0 + 1) * KC) ? ((t0 + 1) * KC < c1 ? (t0 + 1) * KC : c1) : c0;
    s.nH = (int)(head_end - c0);
    s.kcH = kc0;
    s.tH = (int)t0;
    int64_t tail_begin = c1;
    if (c1 > head_end && c1 != (t1 + 1) * KC) tail_begin = t1 * KC > head_end ? t1 * KC : head_end;
    s.nT = (int)(c1 - tail_begin);
    s.tT = (int)t1;
    s.nF = (int)(tail_begin - head_end);
    s.tF = (int)(head_end / KC);
    return s;
}


// per-format launcher, defined (explicitly specialised) in inst_<format>.cu so formats compile in parallel
template <int FAMILY>
cudaError_t mv_launch(const MatvecParams& p, int mb, int grid, int smem, cudaStream_t st);

}  // namespace b200q
