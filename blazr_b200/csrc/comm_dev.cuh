// comm_dev.cuh -- device-side view of the tensor-parallel exchange buffers (one per rank, CUDA IPC mapped into every
// peer; csrc/comm.cu owns them).  The exchange is FUSED into the kernels on both sides of it (SURVEY.md section 5 / 8b
// b200q_matmul_rowpar_allreduce):
//
//   producer  = the row-parallel matvec (o_proj / down_proj).  Its flush paths store every finished row sum -- the exact
//               f64 partial of this rank's K slice -- straight into slot (parity, rank) of EVERY rank's buffer over NVLink
//               (fire-and-forget 8-byte peer stores).  NOTHING ELSE: no fence, no flag.  The all-reduce slots are their own
//               ready flags (the "LL" idea): they are zero between exchanges, an exact zero is published as -0.0, so
//               +0.0 bits mean "not written yet" and an aligned 8-byte store is observed whole or not at all.
//   consumer  = the add+RMSNorm+quantise kernel that follows.  Every thread polls the `world` slot elements of ITS column
//               in its own memory until all are non-zero, writes +0.0 back (slot free for the exchange after next), sums
//               them in RANK ORDER in f64 from +0.0 and rounds once: identical bits on every rank and identical to the
//               1-GPU sum (the partials are exact products accumulated in f64).  An exchange is consumed exactly once.
//   Round 2 measured the flag form (per-CTA fence.acq_rel.sys + arrival atomic, last CTA raises `world` flags, consumer
//   acquires them, then loads) at ~13 us per exchange at TP8 -- a system-scope fence waits for every outstanding NVLink store
//   of the CTA; with self-validating data the exchange costs one NVLink store latency.
//
// The same mechanism carries the lm_head all-gather (column-parallel over the vocabulary): f32 logits are stored into
// region `rank` of every peer's gather area, consumed by the arg-max kernel.  No NCCL, no separate exchange launch.
//
// Buffer layout (bytes):
//   [0, 256)      AR flags  u32 [2 parities][8 source ranks]     written by peers (st.release.sys), read locally
//   [256, 512)    AG flags  u32 [8 source ranks]
//   [512, 1024)   local state: AR epoch, AR done counter, AG epoch, AG done counter (u32 each; never touched by peers)
//   [COMM_HDR_BYTES, +2*world*slot_elems*8)        AR slots  f64 [2][world][slot_elems]
//   [.., + world*gather_elems*4)                   AG region f32 [world][gather_elems]   (initialised to -inf)
// Two AR slot sets alternate by epoch parity: a rank can run at most one exchange ahead of its slowest peer (it needs
// that peer's data of epoch e to finish e), so stores of epoch e+1 never land in a slot a peer is still summing for e, and a
// peer's stores of epoch e+2 (same parity as e) are issued only after it consumed MY epoch e+1 data, which my stream produces
// after my consumer of e has zeroed the slots.  The AR epoch is local state: every CTA of a producer reads it, then arrives
// on the done counter (relaxed, device scope); the last arriver publishes epoch + 1 for the consumer / the next producer.
// The gather region is single-buffered: 2L all-reduces separate two consecutive lm_heads.
#pragma once
#include <stdint.h>

namespace b200q {

constexpr int COMM_MAX_WORLD = 8;
// Consumers poll a slot element at most this many times (one L2 round trip each, ~0.5 us: about a minute) before giving up:
// ranks may be seconds apart at the first exchange after a load, but a lost peer must not hang the GPU for ever.
constexpr int COMM_SPIN_LIMIT = 1 << 27;
constexpr size_t COMM_HDR_BYTES = 8192;
constexpr int COMM_OFF_AR_FLAGS = 0;
constexpr int COMM_OFF_AG_FLAGS = 256;
constexpr int COMM_OFF_AR_EPOCH = 512;
constexpr int COMM_OFF_AR_DONE = 516;
constexpr int COMM_OFF_AG_EPOCH = 520;
constexpr int COMM_OFF_AG_DONE = 524;

struct CommDev {
    uint8_t* peers[COMM_MAX_WORLD];  // base of every rank's buffer (peers[rank] = own); unused entries null
    int rank, world;                 // world == 0: no exchange (plain kernels)
    int64_t slot_elems;              // doubles per (parity, source rank) AR slot
    int64_t gather_elems;            // floats per source-rank gather region
};

enum { RP_NONE = 0, RP_ALLREDUCE = 1, RP_ALLGATHER = 2 };
struct RemoteOut {
    int mode;  // RP_ALLREDUCE / RP_ALLGATHER
    CommDev comm;
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Raise this rank's epoch flag at every rank.  ONE system-scope fence, then `world` fire-and-forget relaxed stores: a
// st.release.sys per peer would serialise `world` NVLink round trips on the last CTA (measured at TP8: the exchange then costs
// more than the matvec it follows).  fence + relaxed store is the release pattern of the PTX memory model.
__device__ __forceinline__ void comm_raise_flags(const struct CommDev& c, int flags_off, int par, unsigned int epoch);
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ size_t comm_ar_slot_off(const CommDev& c, int par, int src) {
    return COMM_HDR_BYTES + ((size_t)(par * c.world + src) * (size_t)c.slot_elems) * sizeof(double);
}
__device__ __forceinline__ size_t comm_ag_off(const CommDev& c, int src) {
    return COMM_HDR_BYTES + (size_t)2 * c.world * (size_t)c.slot_elems * sizeof(double) + (size_t)src * (size_t)c.gather_elems * sizeof(float);
}
__device__ __forceinline__ void comm_raise_flags(const CommDev& c, int flags_off, int par, unsigned int epoch) {
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    for (int r = 0; r < c.world; r++) st_relaxed_sys(reinterpret_cast<unsigned int*>(c.peers[r] + flags_off) + par * COMM_MAX_WORLD + c.rank, epoch);
}
// consumer side: wait until every rank's flag has reached `epoch` (threads 0..world-1 of the CTA poll; caller syncs)
__device__ __forceinline__ void comm_wait_flags(const CommDev& c, int flags_off, int par, unsigned int epoch, int tid) {
    if (tid < c.world) {
        const unsigned int* f = reinterpret_cast<const unsigned int*>(c.peers[c.rank] + flags_off) + par * COMM_MAX_WORLD + tid;
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
        }
    }
}
// ---- self-validating all-reduce slots ----
__device__ __forceinline__ double ar_encode(double v) { return v == 0.0 ? -0.0 : v; }   // +0.0 bits are reserved for "empty"
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// producer bookkeeping (tid 0 of every CTA, after the CTA's last use of the epoch): the last arriver advances the epoch
__device__ __forceinline__ void ar_epoch_arrive(uint8_t* mine, unsigned int epoch, unsigned int n_ctas) {
    unsigned int* done = reinterpret_cast<unsigned int*>(mine + COMM_OFF_AR_DONE);
    unsigned int old;
    asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(done) : "memory");
    if (old == n_ctas - 1u) {
        *done = 0u;                                                                     // next launch (ordered by the stream)
        *reinterpret_cast<unsigned int*>(mine + COMM_OFF_AR_EPOCH) = epoch;
    }
}
// consumer: sum over ranks (rank order, from +0.0) of element idx of the parity-`par` slots; waits for every rank's store and
// leaves the slots empty.  Bounded spin: a lost peer must not hang the GPU (the result is then garbage, never a deadlock).
__device__ __forceinline__ double ar_consume(const CommDev& c, int par, size_t idx) {
    uint8_t* mine = c.peers[c.rank];
    double v[COMM_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < COMM_MAX_WORLD; r++)
        if (r < c.world) v[r] = ld_volatile_f64(reinterpret_cast<const double*>(mine + comm_ar_slot_off(c, par, r)) + idx);
#pragma unroll
    for (int r = 0; r < COMM_MAX_WORLD; r++)
        if (r < c.world) {
            double* q = reinterpret_cast<double*>(mine + comm_ar_slot_off(c, par, r)) + idx;
            for (int spin = 0; __double_as_longlong(v[r]) == 0ll && spin < COMM_SPIN_LIMIT; spin++) v[r] = ld_volatile_f64(q);
            *q = 0.0;
        }
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < COMM_MAX_WORLD; r++)
        if (r < c.world) s += v[r];
    return s;
}
#endif


}  // namespace b200q
