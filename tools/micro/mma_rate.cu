// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of N, operand source of A
// (TMEM ".ts" vs shared memory ".ss") and commit cadence.  One CTA per SM, thread 0 issues.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../blazr_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cstdlib>
#include "gemm_impl.cuh"
using namespace b200q;

__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mode: 0 = ts, 1 = ss.  commit_every: commit to a scratch barrier every c MMAs (0 = never).  nacc: accumulators rotated.
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
// CE: commit to a scratch barrier after every CE MMAs (0 = never).  16 MMAs per unrolled body, compile-time everything.
template <int N, int MODE, int CE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* scratch = bar + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 4);
    uint8_t* btile = smem + 1024;              // 256 rows x 128 B
    uint8_t* atile = smem + 1024 + 256 * 128;  // 128 rows x 128 B
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (256 * 128 + 128 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(btile)[i] = 0;
    if (tid == 0) { mbar_init(bar, 1); mbar_init(scratch, 1); fence_mbar_init(); fence_proxy_async(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0 && elect_one()) {
        const uint64_t bdesc = make_b_desc(smem_u32(btile));
        const uint64_t adesc = make_b_desc(smem_u32(atile));
        const uint32_t a = tmem + 384;
        long long t0 = clock64();
        for (int r = 0; r < reps; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (MODE == 0) tc_mma_ts(tmem, a + (i & 15) * 8, bdesc + (uint64_t)((i & 3) * 2), idesc, 1u);
                else tc_mma_ss(tmem, adesc + (uint64_t)((i & 3) * 2), bdesc + (uint64_t)((i & 3) * 2), idesc, 1u);
                if (CE && ((i + 1) % CE) == 0) tc_commit(scratch);
            }
        }
        tc_commit(bar);
        long long t1 = clock64();
        mbar_wait(bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

template <int N, int MODE, int CE>
void run(long long* out) {
    const int smem = 1024 + 256 * 128 + 128 * 128;
    cudaFuncSetAttribute(rate_kernel<N, MODE, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 64;
    rate_kernel<N, MODE, CE><<<148, 128, smem>>>(reps, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%s N=%3d commit_every=%2d : issue %7.1f total %7.1f cyc/mma %s\n", MODE ? "ss" : "ts", N, CE, h[0] / (reps * 16.0), h[1] / (reps * 16.0),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    run<32, 0, 0>(out); run<32, 0, 4>(out); run<32, 0, 16>(out); run<32, 1, 0>(out); run<32, 1, 4>(out);
    run<64, 0, 0>(out); run<64, 0, 4>(out); run<64, 1, 0>(out);
    run<128, 0, 0>(out); run<128, 0, 4>(out); run<128, 1, 0>(out); run<128, 1, 4>(out);
    run<256, 0, 0>(out); run<256, 0, 4>(out); run<256, 0, 16>(out); run<256, 1, 0>(out); run<256, 1, 4>(out);
    return 0;
}
