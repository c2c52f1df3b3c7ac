// dstep.cu -- host side of the EXPERIMENTAL persistent op-list kernel (dstep_impl.cuh): program builder + launcher.
// Opt-in (blazr_b200/decode.py: B200Q_DSTEP=1); not yet run on hardware -- see DESIGN.md section 8 item 1.
#include <cstring>
#include <new>
#include <vector>

#include "dstep_impl.cuh"

using namespace b200q;

struct b200q_program {
    int device, num_sms, M, finalized;
    std::vector<StepOp> ops;
    StepOp* ops_dev;
    unsigned int* bar_dev;  // [0] arrivals, [32] generation (separate 128-byte lines)
    int nstages, stage_bytes, xhat_bytes, smem_bytes;
    std::vector<void*> owned;  // device scratch owned by the program (argmax candidates)
};

template <int MB>
static cudaError_t launch_dstep(const StepParams& sp, int grid, int smem, cudaStream_t st) {
    static bool configured[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(dstep_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(MV_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, dstep_kernel<MB>, sp);
    if (le != cudaSuccess) return le;
    count_launch();
    return cudaGetLastError();
}

extern "C" {

int32_t b200q_program_create(int32_t device, b200q_program** out) {
    if (!out) return B200Q_ERR_INVALID_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B200Q_ERR_CUDA;
    if (prop.major != 10) return B200Q_ERR_NO_DEVICE;
    b200q_program* p = new (std::nothrow) b200q_program();
    if (!p) return B200Q_ERR_CUDA;
    p->device = device;
    p->num_sms = prop.multiProcessorCount;
    p->M = 0;
    p->finalized = 0;
    p->ops_dev = nullptr;
    p->bar_dev = nullptr;
    *out = p;
    return B200Q_OK;
}

static int32_t set_m(b200q_program* p, int64_t M) {
    if (M != 1 && M != 2 && M != 4) return B200Q_ERR_UNSUPPORTED;   // the record-reading ops take exactly the kernel's MB rows
    if (p->M == 0) p->M = (int)M;
    return p->M == (int)M ? B200Q_OK : B200Q_ERR_INVALID_ARG;
}

int32_t b200q_program_add_normq(b200q_program* p, const float* h_in, const float* delta, float* h_out, const float* norm_w, float eps, int64_t H,
                                int64_t M, void* xq_out) {
    if (!p || p->finalized || !h_in || !h_out || !norm_w || !xq_out || H <= 0 || H % CHUNK_K || (delta && h_in == h_out)) return B200Q_ERR_INVALID_ARG;
    int32_t rc = set_m(p, M);
    if (rc) return rc;
    StepOp op;
    memset(&op, 0, sizeof(op));
    op.type = DS_NORMQ;
    op.M = (int)M; op.H = (int)H; op.h_in = h_in; op.delta = delta; op.h_out = h_out; op.norm_w = norm_w; op.eps = eps; op.xq_out = (uint8_t*)xq_out;
    p->ops.push_back(op);
    return B200Q_OK;
}

int32_t b200q_program_add_swigluq(b200q_program* p, const float* gate_up, int64_t F, int64_t M, void* xq_out) {
    if (!p || p->finalized || !gate_up || !xq_out || F <= 0 || F % 32) return B200Q_ERR_INVALID_ARG;
    int32_t rc = set_m(p, M);
    if (rc) return rc;
    StepOp op;
    memset(&op, 0, sizeof(op));
    op.type = DS_SWIGLUQ;
    op.M = (int)M; op.F = (int)F; op.gate_up = gate_up; op.xq_out = (uint8_t*)xq_out;
    p->ops.push_back(op);
    return B200Q_OK;
}

int32_t b200q_program_add_matvec(b200q_program* p, const b200q_weight* w, const void* xq, int64_t M, void* y, int32_t y_dtype, int64_t ldy,
                                 void* workspace, size_t workspace_bytes) {
    if (!p || p->finalized || !w || !xq || !y || !workspace || ldy < w->N || y_dtype < 0 || y_dtype > 3) return B200Q_ERR_INVALID_ARG;
    if (w->device != p->device || w->perm) return B200Q_ERR_UNSUPPORTED;
    if (w->family != B200Q_FAM_Q4_K && w->family != B200Q_FAM_Q6_K && w->family != B200Q_FAM_Q8_0 && w->family != B200Q_FAM_G4) return B200Q_ERR_UNSUPPORTED;
    int32_t rc = set_m(p, M);
    if (rc) return rc;
    if (workspace_bytes < ((matvec_ws_bytes(w, M) + 255) & ~(size_t)255)) return B200Q_ERR_WORKSPACE;
    StepOp op;
    memset(&op, 0, sizeof(op));
    op.type = DS_MATVEC;
    op.family = w->family; op.gpc = w->gpc; op.chunk_bytes = w->chunk_bytes;
    op.M = (int)M; op.y_dtype = y_dtype; op.KC = (int)w->KC;
    op.N = w->N; op.C = w->T * w->KC; op.ldy = ldy;
    op.w = w->data; op.xq = (const uint8_t*)xq; op.y = y; op.bias = w->bias;
    const size_t cnt = ((size_t)w->T * 4 + 255) & ~(size_t)255;
    op.ws_cnt = reinterpret_cast<unsigned int*>(workspace);
    op.ws_part = reinterpret_cast<double*>((uint8_t*)workspace + cnt);
    p->ops.push_back(op);
    return B200Q_OK;
}

int32_t b200q_program_add_attn(b200q_program* p, const float* qkv, const int32_t* pos, float* cache_k, float* cache_v, const float* rope_table,
                               int32_t n_heads, int32_t n_kv_heads, int32_t head_dim, int32_t max_ctx, int64_t M, void* xq_out) {
    if (!p || p->finalized || !qkv || !pos || !cache_k || !cache_v || !rope_table || !xq_out || n_heads <= 0 || n_kv_heads <= 0 || n_heads % n_kv_heads ||
        max_ctx <= 0)
        return B200Q_ERR_INVALID_ARG;
    if (head_dim != 64 && head_dim != 128) return B200Q_ERR_UNSUPPORTED;
    if (((int64_t)n_heads * head_dim) % CHUNK_K) return B200Q_ERR_INVALID_ARG;
    int32_t rc = set_m(p, M);
    if (rc) return rc;
    StepOp op;
    memset(&op, 0, sizeof(op));
    op.type = DS_ATTN;
    op.M = (int)M; op.qkv = qkv; op.pos = pos; op.cache_k = cache_k; op.cache_v = cache_v; op.rope = rope_table;
    op.nh = n_heads; op.nkv = n_kv_heads; op.hd = head_dim; op.max_ctx = max_ctx; op.xq_out = (uint8_t*)xq_out;
    p->ops.push_back(op);
    return B200Q_OK;
}

/* greedy sampling inside the program: two ops (per-CTA candidates, then the final pick by CTA 0) */
int32_t b200q_program_add_argmax(b200q_program* p, const float* logits, int64_t V, int64_t M, int64_t* out_ids, int32_t* pos_inc) {
    if (!p || p->finalized || !logits || !out_ids || V <= 0) return B200Q_ERR_INVALID_ARG;
    int32_t rc = set_m(p, M);
    if (rc) return rc;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    float* cv = nullptr;
    int* ci = nullptr;
    cudaError_t e = cudaMalloc(&cv, sizeof(float) * (size_t)M * p->num_sms);
    if (e == cudaSuccess) e = cudaMalloc(&ci, sizeof(int) * (size_t)M * p->num_sms);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return B200Q_ERR_CUDA;
    p->owned.push_back(cv);
    p->owned.push_back(ci);
    StepOp op;
    memset(&op, 0, sizeof(op));
    op.type = DS_ARGMAX_A;
    op.M = (int)M; op.V = (int)V; op.logits = logits; op.cand_val = cv; op.cand_idx = ci; op.ids = out_ids; op.pos_inc = pos_inc;
    p->ops.push_back(op);
    op.type = DS_ARGMAX_B;
    p->ops.push_back(op);
    return B200Q_OK;
}

int32_t b200q_program_add_embed(b200q_program* p, const void* table_f16, const int64_t* ids, int64_t H, int64_t M, float* h) {
    if (!p || p->finalized || !table_f16 || !ids || !h || H <= 0) return B200Q_ERR_INVALID_ARG;
    int32_t rc = set_m(p, M);
    if (rc) return rc;
    StepOp op;
    memset(&op, 0, sizeof(op));
    op.type = DS_EMBED;
    op.M = (int)M; op.H = (int)H; op.table = (const __half*)table_f16; op.ids = const_cast<int64_t*>(ids); op.h_out = h;
    p->ops.push_back(op);
    return B200Q_OK;
}

int32_t b200q_program_finalize(b200q_program* p) {
    if (!p || p->finalized || p->ops.empty()) return B200Q_ERR_INVALID_ARG;
    int max_chunk = 0;
    size_t max_x = 0;
    for (const StepOp& op : p->ops)
        if (op.type == DS_MATVEC) {
            if (op.chunk_bytes > max_chunk) max_chunk = op.chunk_bytes;
            const size_t x = (size_t)op.KC * op.M * ACT_REC_BYTES;
            if (x > max_x) max_x = x;
        } else if (op.type == DS_ATTN) {
            // s_o [512 / (hd/4)][hd] f64 + 16 f64 + 16 f32 + 3 hd f32 + max_ctx f32 (dstep_impl.cuh ds_attention)
            const size_t x = (size_t)(DS_CONSUMERS / (op.hd / 4)) * op.hd * 8 + 128 + 64 + 3 * (size_t)op.hd * 4 + (size_t)op.max_ctx * 4;
            if (x > max_x) max_x = x;
        } else if (op.type == DS_ARGMAX_A) {
            if (max_x < 256) max_x = 256;
        }
    p->stage_bytes = max_chunk > 0 ? ((max_chunk + 127) & ~127) : 128;
    p->xhat_bytes = (int)((max_x + 127) & ~(size_t)127);
    const int budget = 220 * 1024 - 1024 - p->xhat_bytes;
    int nst = budget / p->stage_bytes;
    if (nst > MV_MAX_STAGES) nst = MV_MAX_STAGES;
    if (nst < 2) return B200Q_ERR_UNSUPPORTED;
    p->nstages = nst;
    p->smem_bytes = 1024 + p->xhat_bytes + nst * p->stage_bytes;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    cudaError_t e = cudaMalloc(&p->ops_dev, sizeof(StepOp) * p->ops.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->ops_dev, p->ops.data(), sizeof(StepOp) * p->ops.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&p->bar_dev, 256);
    if (e == cudaSuccess) e = cudaMemset(p->bar_dev, 0, 256);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return B200Q_ERR_CUDA;
    p->finalized = 1;
    return B200Q_OK;
}

int32_t b200q_program_launch(const b200q_program* p, void* stream) {
    if (!p || !p->finalized) return B200Q_ERR_INVALID_ARG;
    StepParams sp;
    sp.ops = p->ops_dev;
    sp.n_ops = (int)p->ops.size();
    sp.nstages = p->nstages;
    sp.stage_bytes = p->stage_bytes;
    sp.xhat_bytes = p->xhat_bytes;
    sp.bar_count = p->bar_dev;
    sp.bar_gen = p->bar_dev + 32;
    cudaError_t e;
    switch (p->M) {
        case 1: e = launch_dstep<1>(sp, p->num_sms, p->smem_bytes, (cudaStream_t)stream); break;
        case 2: e = launch_dstep<2>(sp, p->num_sms, p->smem_bytes, (cudaStream_t)stream); break;
        default: e = launch_dstep<4>(sp, p->num_sms, p->smem_bytes, (cudaStream_t)stream); break;
    }
    return e == cudaSuccess ? B200Q_OK : B200Q_ERR_CUDA;
}

int32_t b200q_program_free(b200q_program* p) {
    if (!p) return B200Q_OK;
    if (p->ops_dev) cudaFree(p->ops_dev);
    if (p->bar_dev) cudaFree(p->bar_dev);
    for (void* q : p->owned) cudaFree(q);
    delete p;
    return B200Q_OK;
}

}  // extern "C"
