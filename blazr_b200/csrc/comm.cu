// comm.cu -- tensor-parallel exchange over NVLink peer memory (one process per GPU, CUDA IPC), no NCCL on the
// decode path.  The row-parallel projections (o_proj, down_proj: reference src/engine/tensor_parallel.rs, SURVEY
// section 8e) need one all-reduce(sum) of [M, hidden] per call: 32-256 KB, i.e. pure latency.  One-shot "push"
// all-reduce: every rank stores its f64 partial sums straight into a slot of every peer's receive buffer
// (fire-and-forget NVLink stores), raises a per-CTA epoch flag at the peer with a system-scope release, waits for
// the peers' flags in its own memory and sums the `world` slots in RANK ORDER in f64 -> every rank computes the
// same bits, and because the partials are the matvec's exact-product f64 accumulators the result equals the
// 1-GPU output bit for bit (up to the same rare double-rounding cases), independent of the TP degree.
// Two slot sets alternate by epoch parity: a rank can run at most one all-reduce ahead of its slowest peer
// (it needs that peer's flag of epoch e to finish e), so writes of epoch e+1 never land in a slot still being read.
#include "common.cuh"
#include "internal.h"
#include <cstring>

namespace b200q {

constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_MAX_CTAS = 32;
constexpr size_t COMM_HDR_BYTES = 8192;  // flags [2][8][32] u32 (2 KB) + epochs [32] u32, padded

struct CommDev {
    uint8_t* peers[COMM_MAX_WORLD];  // base of every rank's buffer (peers[rank] = own)
    int rank, world;
    int64_t slot_elems;              // doubles per (parity, source rank) slot
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) allreduce_push_kernel(const CommDev c, const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();  // src is written by the row-parallel matvec just ahead in the stream
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
    uint8_t* mine = c.peers[c.rank];
    unsigned int* epochs = reinterpret_cast<unsigned int*>(mine + 2 * COMM_MAX_WORLD * COMM_MAX_CTAS * 4);
    __shared__ unsigned int s_epoch;
    if (tid == 0) s_epoch = ++epochs[cta];
    __syncthreads();
    const unsigned int epoch = s_epoch;
    const int par = (int)(epoch & 1u);
    // slice of this CTA, in double2 units
    const int64_t n2 = (n + 1) / 2;
    const int64_t per = (n2 + G - 1) / G;
    const int64_t i0 = (int64_t)cta * per, i1 = min(n2, i0 + per);
    // 1. push my partials into slot (par, rank) of every rank (own buffer included)
    for (int64_t i = i0 + tid; i < i1; i += blockDim.x) {
        double2 v;
        v.x = src[2 * i];
        v.y = (2 * i + 1 < n) ? src[2 * i + 1] : 0.0;
        for (int p = 0; p < c.world; p++) {
            double2* slot = reinterpret_cast<double2*>(c.peers[p] + COMM_HDR_BYTES) + ((size_t)(par * c.world + c.rank) * c.slot_elems) / 2;
            slot[i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag at every peer, then wait for every peer's flag here
    if (tid < c.world) {
        unsigned int* flag = reinterpret_cast<unsigned int*>(c.peers[tid]) + ((size_t)(par * COMM_MAX_WORLD + c.rank) * COMM_MAX_CTAS + cta);
        st_release_sys(flag, epoch);
        const unsigned int* wait = reinterpret_cast<const unsigned int*>(mine) + ((size_t)(par * COMM_MAX_WORLD + tid) * COMM_MAX_CTAS + cta);
        while ((int)(ld_acquire_sys(wait) - epoch) < 0) {
        }
    }
    __syncthreads();
    // 3. sum the slots in rank order (identical on every rank) and round once to f32
    const double2* slots = reinterpret_cast<const double2*>(mine + COMM_HDR_BYTES);
    for (int64_t i = i0 + tid; i < i1; i += blockDim.x) {
        double2 s = make_double2(0.0, 0.0);
        for (int r = 0; r < c.world; r++) {
            const double2 v = __ldcg(slots + ((size_t)(par * c.world + r) * c.slot_elems) / 2 + i);
            s.x += v.x;
            s.y += v.y;
        }
        dst[2 * i] = (float)s.x;
        if (2 * i + 1 < n) dst[2 * i + 1] = (float)s.y;
    }
}

}  // namespace b200q

using namespace b200q;

struct b200q_comm {
    int rank, world, device;
    int64_t slot_elems;
    size_t bytes;
    uint8_t* local;
    uint8_t* peers[COMM_MAX_WORLD];
    bool opened[COMM_MAX_WORLD];
};

extern "C" {

int32_t b200q_comm_create(int32_t rank, int32_t world, int64_t max_elems, int32_t device, b200q_comm** out) {
    if (!out || world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world || max_elems < 1) return B200Q_ERR_INVALID_ARG;
    b200q_comm* c = new b200q_comm();
    c->rank = rank; c->world = world; c->device = device;
    c->slot_elems = (max_elems + 1) / 2 * 2;
    c->bytes = COMM_HDR_BYTES + (size_t)2 * world * c->slot_elems * sizeof(double);
    for (int i = 0; i < COMM_MAX_WORLD; i++) { c->peers[i] = nullptr; c->opened[i] = false; }
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    cudaError_t e = cudaMalloc(&c->local, c->bytes);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaSetDevice(prev);
    if (e != cudaSuccess) { delete c; return B200Q_ERR_CUDA; }
    c->peers[rank] = c->local;
    *out = c;
    return B200Q_OK;
}

/* 64-byte cudaIpcMemHandle_t of this rank's buffer: all-gather these (torch.distributed) and pass them to connect */
int32_t b200q_comm_handle(const b200q_comm* c, void* out64) {
    if (!c || !out64) return B200Q_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, c->local) != cudaSuccess) return B200Q_ERR_CUDA;
    memcpy(out64, &h, sizeof(h));
    return B200Q_OK;
}

int32_t b200q_comm_connect(b200q_comm* c, const void* handles) {
    if (!c || !handles) return B200Q_ERR_INVALID_ARG;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    int32_t rc = B200Q_OK;
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)handles + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { rc = B200Q_ERR_CUDA; break; }
        c->peers[r] = (uint8_t*)p;
        c->opened[r] = true;
    }
    cudaSetDevice(prev);
    return rc;
}

int32_t b200q_comm_free(b200q_comm* c) {
    if (!c) return B200Q_OK;
    for (int r = 0; r < c->world; r++)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peers[r]);
    if (c->local) cudaFree(c->local);
    delete c;
    return B200Q_OK;
}

/* dst[n] (f32) = sum over ranks of src[n] (f64 partial sums), summed in rank order: identical bits on every rank.
 * Graph-capturable: no host synchronisation, epochs live in device memory. */
int32_t b200q_allreduce_f64(b200q_comm* c, const double* src, float* dst, int64_t n, void* stream) {
    if (!c || !src || !dst || n < 1 || n > c->slot_elems) return B200Q_ERR_INVALID_ARG;
    for (int r = 0; r < c->world; r++)
        if (!c->peers[r]) return B200Q_ERR_INVALID_ARG;
    CommDev d;
    for (int r = 0; r < COMM_MAX_WORLD; r++) d.peers[r] = c->peers[r];
    d.rank = c->rank; d.world = c->world; d.slot_elems = c->slot_elems;
    int grid = (int)((n + 2047) / 2048);  // >= 1024 double2 per CTA
    if (grid > COMM_MAX_CTAS) grid = COMM_MAX_CTAS;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, allreduce_push_kernel, d, src, dst, n);
    count_launch();
    return e == cudaSuccess ? B200Q_OK : B200Q_ERR_CUDA;
}

}  // extern "C"
