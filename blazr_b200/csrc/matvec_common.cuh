// matvec_common.cuh -- parameters and stream-K helpers shared by the matvec kernel and its host dispatcher.
#pragma once
#include "comm_dev.cuh"
#include "formats.cuh"
#include "internal.h"
#include "streamk_plan.cuh"

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
    // successor hint (b200q_weight_set_next): once this CTA's own chunks are all requested, its producer asks the TMA
    // engine to pull the first next_pf chunks of every successor CTA range (in that CTA's processing order) into L2, so
    // HBM keeps streaming through this launch's tail, the kernel boundary and whatever glue operators run in between
    const uint8_t* next_w;
    int64_t next_C, next_KC;
    int next_chunk_bytes, next_G, next_pf;
    // fused tensor-parallel exchange (comm_dev.cuh): RP_ALLREDUCE = finished row sums (f64) go to slot (parity, rank) of every
    // rank's buffer instead of y; RP_ALLGATHER = f32 outputs go to region `rank` of every rank's gather area.  Index inside
    // the slot / region = m * ldy + n.  The last CTA of the launch raises this rank's epoch flag at every peer.
    int rp_mode;
    CommDev comm;
    // fused SwiGLU epilogue (OUT_SWIGLU): w is a gate|up weight whose rows are interleaved per 128-row tile (tile t, row
    // 8 w + 4 s + g  <->  (s == 0 ? gate : up)[64 t + 4 w + g]); every finished tile yields 64 values silu(gate) * up =
    // two 32-blocks, quantised straight into the next matvec's activation records xq_out[kc][epi_rows][320].
    uint8_t* xq_out;
    int epi_F;      // rows of gate (= rows of up) = N / 2
    int epi_rows;   // record rows per k-chunk: M (plain) or n_slots (grouped)
    int plan32;     // (C + 1) * G < 2^32: the stream-K plan is computed with 32-bit divisions
    // dual-format launch (kernel template F2 != F; b200q_weight_set_pair): a second weight of another format whose output rows
    // directly follow the first one's (y2 = y + N1, N1 % 128 == 0) rides in the same stream-K grid -- tiles [0, T1) belong to w
    // (format F, cb1 bytes per chunk), tiles [T1, T1 + T2) to w2 (format F2, cb2 bytes per chunk).  chunk_bytes is then the
    // larger of the two: the offset of the activation records inside a ring stage.  One launch instead of two for the q|k + v
    // projections of Q4_K_M files (V is Q6_K in about half of the layers).
    const uint8_t* w2;
    int cb1, cb2, T1, gpc2;
    const uint8_t* const* w_table;
    const int32_t* sel;
    int n_experts;        // grouped: entries of w_table; a selection outside [0, n_experts) is treated as "not hosted" (slot skipped)
    int tpw;              // tiles per weight
    int x_rows, x_slot_div;
    int64_t y_slot_stride;
};

// per-format launcher, defined (explicitly specialised) in inst_<format>.cu so formats compile in parallel
template <int FAMILY>
cudaError_t mv_launch(const MatvecParams& p, int mb, int grid, int smem, cudaStream_t st);
// dual-format launcher (inst_q4k_q6k.cu): first weight Q4_K, second Q6_K
cudaError_t mv_launch_dual_q4k_q6k(const MatvecParams& p, int mb, int grid, int smem, cudaStream_t st);

}  // namespace b200q
