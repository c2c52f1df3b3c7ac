// streamk_plan.cuh -- the stream-K work split of the matvec kernel: which chunks a CTA owns, in which order it walks
// them (head / tail / full segments) and who contributes to a split tile.  Pure integer logic with no CUDA
// dependency, so tests/test_streamk_plan.py compiles it for the host and checks it exhaustively on the CPU.
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif

namespace b200q {

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

// 32-bit forms (the host guarantees (C + 1) * G < 2^32): a 64-bit division is a ~150-instruction subroutine on the GPU and
// the plan sits on the critical path of every launch's prologue
__device__ __forceinline__ uint32_t sk_begin32(uint32_t g, uint32_t C, uint32_t G) { return g * C / G; }
__device__ __forceinline__ SkPlan sk_plan32(uint32_t c0, uint32_t c1, uint32_t KC) {
    SkPlan s;
    const uint32_t t0 = c0 / KC, t1 = (c1 - 1) / KC;
    const uint32_t kc0 = c0 - t0 * KC;
    const uint32_t head_end = (kc0 != 0 || c1 < (t0 + 1) * KC) ? ((t0 + 1) * KC < c1 ? (t0 + 1) * KC : c1) : c0;
    s.nH = (int)(head_end - c0);
    s.kcH = (int)kc0;
    s.tH = (int)t0;
    uint32_t tail_begin = c1;
    if (c1 > head_end && c1 != (t1 + 1) * KC) tail_begin = t1 * KC > head_end ? t1 * KC : head_end;
    s.nT = (int)(c1 - tail_begin);
    s.tT = (int)t1;
    s.nF = (int)(tail_begin - head_end);
    s.tF = (int)(head_end / KC);
    return s;
}

// index (in the tile-major chunk array) of the j-th chunk a CTA with range [c0, ...) and plan sp processes
__device__ __forceinline__ int64_t sk_chunk_at(const SkPlan& sp, int64_t c0, int j) {
    if (j < sp.nH) return c0 + j;
    if (j < sp.nH + sp.nT) return c0 + sp.nH + sp.nF + (j - sp.nH);
    return c0 + sp.nH + (j - sp.nH - sp.nT);
}

}  // namespace b200q
