// host_smoke.cpp -- exercises the C++ host mirror end to end on one GPU: Q8_0 weight upload, dequantize,
// quant_matmul (M = 1), error codes.  Exit code 0 = ok.  (The numeric parity suite lives in tests/; this only
// proves the compiled-language host side drives the C ABI.)
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "b200q.hpp"

using namespace b200q_host;

static uint16_t f2h(float f) {  // finite, normal range only (test data)
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000, exp = ((x >> 23) & 0xFF) - 112, man = (x >> 13) & 0x3FF;
    return (uint16_t)(sign | (exp << 10) | man);
}
static float h2f(uint16_t h) {
    uint32_t x = ((uint32_t)(h & 0x8000) << 16) | ((((h >> 10) & 0x1F) + 112) << 23) | ((uint32_t)(h & 0x3FF) << 13);
    float f;
    memcpy(&f, &x, 4);
    return f;
}

int main() {
    try {
        const int64_t N = 256, K = 512;
        std::vector<uint8_t> blocks(N * (K / 32) * 34);
        std::vector<float> wref(N * K);
        uint32_t s = 12345;
        for (int64_t b = 0; b < N * (K / 32); b++) {
            uint16_t d = f2h(0.01f + 0.001f * (b % 7));
            memcpy(&blocks[b * 34], &d, 2);
            for (int j = 0; j < 32; j++) {
                s = s * 1664525u + 1013904223u;
                int8_t q = (int8_t)(s >> 24);
                blocks[b * 34 + 2 + j] = (uint8_t)q;
                wref[b * 32 + j] = h2f(d) * (float)q;
            }
        }
        B200Client client(0, nullptr);
        QuantTensor w = client.weight_from_ggml(8, blocks.data(), false, N, K);
        // DequantOps
        float* dq;
        cudaMalloc(&dq, N * K * 4);
        client.dequantize(w, dq, DType::F32);
        std::vector<float> got(N * K);
        cudaMemcpy(got.data(), dq, N * K * 4, cudaMemcpyDeviceToHost);
        for (int64_t i = 0; i < N * K; i++)
            if (got[i] != wref[i]) { printf("dequant mismatch at %lld\n", (long long)i); return 1; }
        // QuantMatmulOps, M = 1
        std::vector<float> x(K);
        for (int64_t k = 0; k < K; k++) x[k] = sinf(0.37f * k);
        float *dx, *dy;
        void* ws;
        size_t wsb = client.workspace_bytes(w, 1);
        cudaMalloc(&dx, K * 4); cudaMalloc(&dy, N * 4); cudaMalloc(&ws, wsb);
        cudaMemset(ws, 0, wsb);
        cudaMemcpy(dx, x.data(), K * 4, cudaMemcpyHostToDevice);
        client.quant_matmul(dx, DType::F32, 1, K, w, dy, DType::F32, N, ws, wsb);
        std::vector<float> y(N);
        cudaMemcpy(y.data(), dy, N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int64_t n = 0; n < N; n++) {
            double acc = 0;
            for (int64_t k = 0; k < K; k++) acc += (double)wref[n * K + k] * x[k];
            maxerr = std::fmax(maxerr, std::fabs(acc - y[n]));
            maxref = std::fmax(maxref, std::fabs(acc));
        }
        if (maxerr / maxref > 1e-2) { printf("matmul rel err %g\n", maxerr / maxref); return 1; }
        // errors are codes, not crashes
        // (ggml type 9 = Q8_1 is an activation-side format: it has no weight kernel and never will)
        try {
            client.weight_from_ggml(9, blocks.data(), false, 1, 256);
            printf("weight_from_ggml(type 9 / Q8_1) succeeded, expected B200Q_ERR_UNSUPPORTED\n");
            return 1;
        } catch (const BackendError& e) {
            if (e.code != B200Q_ERR_UNSUPPORTED) { printf("weight_from_ggml(type 9): error code %d, expected B200Q_ERR_UNSUPPORTED (%d)\n", (int)e.code, (int)B200Q_ERR_UNSUPPORTED); return 1; }
        }
        TensorParallelState tp{2, 3};
        auto r = tp.shard_range(10);
        if (r.first != 7 || r.second != 10) { printf("shard_range(10) at rank 2 of 3 = [%lld, %lld), expected [7, 10)\n", (long long)r.first, (long long)r.second); return 1; }  // reference tensor_parallel.rs:186
        printf("host_smoke ok: dequant bit-exact, matmul rel err %.2e\n", maxerr / maxref);
        return 0;
    } catch (const std::exception& e) {
        printf("FAILED: %s\n", e.what());
        return 2;
    }
}
