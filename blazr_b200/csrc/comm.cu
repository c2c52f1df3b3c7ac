// comm.cu -- tensor-parallel exchange over NVLink peer memory (one process per GPU, CUDA IPC), no NCCL on the decode
// path.  Buffer layout, protocol and the memory-ordering argument are in comm_dev.cuh.  Three users of the buffers:
//   * FUSED (the decode path): b200q_matmul_q8_rowpar / b200q_matmul_q8_gather make the matvec kernel itself push its
//     finished row sums to every peer (csrc/matvec_impl.cuh), b200q_allreduce_add_rmsnorm_quant / b200q_argmax_gathered
//     (csrc/decode_ops.cu) are the consumers -- the exchange costs no launch of its own;
//   * b200q_allreduce_finish: stand-alone consumer (reduced f32 vector) for callers that need the sum itself;
//   * b200q_allreduce: stand-alone one-shot push all-reduce of an f32 / f64 device vector (expert-parallel partial
//     outputs, tests), same slots / flags / epochs, so it can be mixed with the fused form in stream order.
// Reference: the NCCL all-reduce blazr's row-parallel linears issue (src/engine/tensor_parallel.rs:125-163).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "comm_dev.cuh"
#include "common.cuh"
#include "internal.h"

namespace b200q {

// implemented in api.cu: sets the thread-local b200q_last_error message
int32_t set_error(int32_t code, const char* fmt, ...);

constexpr int AR_MAX_CTAS = 32;

// stand-alone all-reduce: push (self-validating slots) -> every thread waits for all ranks' values of its elements -> ordered f64 sum
template <typename T>
__global__ void __launch_bounds__(256) allreduce_kernel(const CommDev c, const T* __restrict__ src, float* __restrict__ dst, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();  // src is written by the kernel just ahead in the stream
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
    uint8_t* mine = c.peers[c.rank];
    const unsigned int epoch = __ldcg(reinterpret_cast<const unsigned int*>(mine + COMM_OFF_AR_EPOCH)) + 1u;
    const int par = (int)(epoch & 1u);
    const int64_t per = (n + G - 1) / G;
    const int64_t i0 = (int64_t)cta * per, i1 = min(n, i0 + per);
    const size_t off = comm_ar_slot_off(c, par, c.rank);
    for (int64_t i = i0 + tid; i < i1; i += blockDim.x) {
        const double v = ar_encode((double)src[i]);
        for (int p = 0; p < c.world; p++) reinterpret_cast<double*>(c.peers[p] + off)[i] = v;
    }
    __syncthreads();  // every thread of the CTA has read the epoch
    if (tid == 0) ar_epoch_arrive(mine, epoch, (unsigned int)G);
    for (int64_t i = i0 + tid; i < i1; i += blockDim.x) dst[i] = (float)ar_consume(c, par, (size_t)i);
}

// stand-alone consumer of a fused exchange: dst = f32(sum over ranks of slot[r]) in rank order
__global__ void __launch_bounds__(256) allreduce_finish_kernel(const CommDev c, float* __restrict__ dst, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();  // the producing matvec of THIS rank has completed: its epoch is published
    const int tid = threadIdx.x;
    const uint8_t* mine = c.peers[c.rank];
    const unsigned int epoch = __ldcg(reinterpret_cast<const unsigned int*>(mine + COMM_OFF_AR_EPOCH));
    const int par = (int)(epoch & 1u);
    (void)tid;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = (float)ar_consume(c, par, (size_t)i);
}

__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace b200q

using namespace b200q;

struct b200q_comm {
    int rank, world, device;
    int64_t slot_elems, gather_elems;
    size_t bytes;
    uint8_t* local;
    uint8_t* peers[COMM_MAX_WORLD];
    bool opened[COMM_MAX_WORLD];
};

namespace b200q {
bool comm_dev(const b200q_comm* c, CommDev* d) {
    for (int r = 0; r < COMM_MAX_WORLD; r++) d->peers[r] = r < c->world ? c->peers[r] : nullptr;
    d->rank = c->rank;
    d->world = c->world;
    d->slot_elems = c->slot_elems;
    d->gather_elems = c->gather_elems;
    for (int r = 0; r < c->world; r++)
        if (!c->peers[r]) return false;
    return true;
}
int64_t comm_slot_elems(const b200q_comm* c) { return c->slot_elems; }
int64_t comm_gather_elems(const b200q_comm* c) { return c->gather_elems; }
int comm_device(const b200q_comm* c) { return c->device; }
}  // namespace b200q

template <typename... Args>
static cudaError_t launch_pdl_comm(void (*kern)(Args...), int grid, int block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    count_launch();
    return e;
}

extern "C" {

int32_t b200q_comm_create(int32_t rank, int32_t world, int64_t max_elems, int64_t gather_elems, int32_t device, b200q_comm** out) {
    if (!out || world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world || max_elems < 1 || gather_elems < 0)
        return set_error(B200Q_ERR_INVALID_ARG, "comm_create: rank %d / world %d (max %d), max_elems %lld, gather_elems %lld", rank, world, COMM_MAX_WORLD,
                         (long long)max_elems, (long long)gather_elems);
    b200q_comm* c = new b200q_comm();
    c->rank = rank; c->world = world; c->device = device;
    c->slot_elems = (max_elems + 1) / 2 * 2;
    c->gather_elems = (gather_elems + 3) / 4 * 4;
    c->bytes = COMM_HDR_BYTES + (size_t)2 * world * c->slot_elems * sizeof(double) + (size_t)world * c->gather_elems * sizeof(float);
    for (int i = 0; i < COMM_MAX_WORLD; i++) { c->peers[i] = nullptr; c->opened[i] = false; }
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    cudaError_t e = cudaMalloc(&c->local, c->bytes);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->bytes);
    if (e == cudaSuccess && c->gather_elems > 0) {  // columns no rank ever writes (vocabulary padding) must lose every arg-max
        float* gp = reinterpret_cast<float*>(c->local + COMM_HDR_BYTES + (size_t)2 * world * c->slot_elems * sizeof(double));
        fill_f32_kernel<<<64, 256>>>(gp, (int64_t)world * c->gather_elems, -INFINITY);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        delete c;
        return set_error(B200Q_ERR_CUDA, "comm_create: %s", cudaGetErrorString(e));
    }
    c->peers[rank] = c->local;
    *out = c;
    return B200Q_OK;
}

/* 64-byte cudaIpcMemHandle_t of this rank's buffer: all-gather these (torch.distributed) and pass them to connect */
int32_t b200q_comm_handle(const b200q_comm* c, void* out64) {
    if (!c || !out64) return set_error(B200Q_ERR_INVALID_ARG, "comm_handle: null argument");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, c->local);
    if (e != cudaSuccess) return set_error(B200Q_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    memcpy(out64, &h, sizeof(h));
    return B200Q_OK;
}

int32_t b200q_comm_connect(b200q_comm* c, const void* handles) {
    if (!c || !handles) return set_error(B200Q_ERR_INVALID_ARG, "comm_connect: null argument");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    int32_t rc = B200Q_OK;
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)handles + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { rc = set_error(B200Q_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e)); break; }
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

/* device pointer of this rank's gather area: f32 [world][gather_elems] (region r = what rank r stored) */
int32_t b200q_comm_gather_ptr(const b200q_comm* c, void** ptr, int64_t* elems_per_rank) {
    if (!c || !ptr) return set_error(B200Q_ERR_INVALID_ARG, "comm_gather_ptr: null argument");
    *ptr = c->local + COMM_HDR_BYTES + (size_t)2 * c->world * c->slot_elems * sizeof(double);
    if (elems_per_rank) *elems_per_rank = c->gather_elems;
    return B200Q_OK;
}

/* dst[n] (f32) = sum over ranks of src[n] (src_dtype B200Q_F32 or B200Q_F64), summed in f64 in rank order: identical bits on
 * every rank.  Graph-capturable: no host synchronisation, epochs live in device memory. */
int32_t b200q_allreduce(b200q_comm* c, const void* src, int32_t src_dtype, float* dst, int64_t n, void* stream) {
    if (!c || !src || !dst || n < 1 || n > c->slot_elems) return set_error(B200Q_ERR_INVALID_ARG, "allreduce: n = %lld outside [1, %lld]", (long long)n, c ? (long long)c->slot_elems : 0ll);
    if (src_dtype != B200Q_F32 && src_dtype != B200Q_F64) return set_error(B200Q_ERR_INVALID_ARG, "allreduce: src_dtype must be f32 or f64");
    CommDev d;
    if (!comm_dev(c, &d)) return set_error(B200Q_ERR_INVALID_ARG, "allreduce: communicator not connected");
    int grid = (int)((n + 2047) / 2048);
    if (grid > AR_MAX_CTAS) grid = AR_MAX_CTAS;
    cudaError_t e = src_dtype == B200Q_F64 ? launch_pdl_comm(allreduce_kernel<double>, grid, 256, (cudaStream_t)stream, (const CommDev)d, (const double*)src, dst, n)
                                           : launch_pdl_comm(allreduce_kernel<float>, grid, 256, (cudaStream_t)stream, (const CommDev)d, (const float*)src, dst, n);
    return e == cudaSuccess ? B200Q_OK : set_error(B200Q_ERR_CUDA, "allreduce launch: %s", cudaGetErrorString(e));
}

int32_t b200q_allreduce_f64(b200q_comm* c, const double* src, float* dst, int64_t n, void* stream) {
    return b200q_allreduce(c, src, B200Q_F64, dst, n, stream);
}

/* consumer of b200q_matmul_q8_rowpar for callers that want the reduced vector itself */
int32_t b200q_allreduce_finish(b200q_comm* c, float* dst, int64_t n, void* stream) {
    if (!c || !dst || n < 1 || n > c->slot_elems) return set_error(B200Q_ERR_INVALID_ARG, "allreduce_finish: n = %lld outside [1, %lld]", (long long)n, c ? (long long)c->slot_elems : 0ll);
    CommDev d;
    if (!comm_dev(c, &d)) return set_error(B200Q_ERR_INVALID_ARG, "allreduce_finish: communicator not connected");
    int grid = (int)((n + 255) / 256);   // one element per thread up to 8192: one round of `world` loads in flight per thread
    if (grid > AR_MAX_CTAS) grid = AR_MAX_CTAS;
    cudaError_t e = launch_pdl_comm(allreduce_finish_kernel, grid, 256, (cudaStream_t)stream, (const CommDev)d, dst, n);
    return e == cudaSuccess ? B200Q_OK : set_error(B200Q_ERR_CUDA, "allreduce_finish launch: %s", cudaGetErrorString(e));
}

}  // extern "C"
