// comm_dev.cuh -- device-side view of the tensor-parallel exchange buffers (one per rank, CUDA IPC mapped into every
// peer; csrc/comm.cu owns them).  The exchange is FUSED into the kernels on both sides of it (SURVEY.md section 5 / 8b
// b200q_matmul_rowpar_allreduce):
//
//   producer  = the row-parallel matvec (o_proj / down_proj).  Its flush paths store every finished row sum -- the exact
//               f64 partial of this rank's K slice -- straight into slot (parity, rank) of EVERY rank's buffer over NVLink
//               (fire-and-forget peer stores); the last CTA of the launch (device-scope arrival counter after a system
//               fence) raises this rank's epoch flag at every peer with st.release.sys.
//   consumer  = the add+RMSNorm+quantise kernel that follows.  It polls the `world` flags in its OWN memory
//               (ld.acquire.sys), sums the slots in RANK ORDER in f64 and rounds once: identical bits on every rank and
//               identical to the 1-GPU sum (the partials are exact products accumulated in f64).
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
// that peer's flag of epoch e to finish e), so stores of epoch e+1 never land in a slot a peer is still summing for e.
// The gather region is single-buffered: 2L all-reduces separate two consecutive lm_heads.
#pragma once
#include <stdint.h>

namespace b200q {

constexpr int COMM_MAX_WORLD = 8;
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
#endif

}  // namespace b200q
