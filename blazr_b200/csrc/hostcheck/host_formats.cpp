// host_formats.cpp -- CPU build of the device tile-layout code (formats.cuh) for tests/test_host_formats.py.
// For every format: repack_row (canonical ggml blocks -> chunk) followed by load_unit (chunk -> 32 integer weights +
// affine map) must reproduce the oracle's decomposition (q = v - off, a, b per sub-block) bit for bit.
// Test infrastructure only (built into blazr_b200/csrc/hostcheck/libhostformats.so by the test).
#define B200Q_HOST_CHECK 1
#include "../formats.cuh"

#include <vector>

using namespace b200q;

template <class F>
static int dump(const uint8_t* blocks, int64_t N, int64_t K, int gpc, int8_t* q, float* a, float* b) {
    const int BE = F::src_block_elems(), BB = F::src_block_bytes();
    const int64_t row_bytes = K / BE * BB;
    const int64_t KC = (K + CHUNK_K - 1) / CHUNK_K, T = (N + TILE_ROWS - 1) / TILE_ROWS;
    const FmtMeta meta{gpc};
    std::vector<uint8_t> chunk((size_t)F::chunk_bytes(gpc) + 64);
    const int nsub = CHUNK_K / F::SUB;
    for (int64_t t = 0; t < T; t++)
        for (int64_t kc = 0; kc < KC; kc++) {
            for (int r = 0; r < TILE_ROWS; r++) {
                const int64_t n = t * TILE_ROWS + r;
                int nvalid = 0;
                if (n < N) {
                    const int64_t rem = (K - kc * CHUNK_K) / BE;
                    const int per = CHUNK_K / BE;
                    nvalid = (int)(rem < per ? rem : per);
                }
                const uint8_t* s = blocks + (n < N ? n : 0) * row_bytes + (kc * CHUNK_K / BE) * BB;
                F::repack_row(s, nvalid, chunk.data(), r, meta);
            }
            for (int r = 0; r < TILE_ROWS; r++) {
                const int64_t n = t * TILE_ROWS + r;
                if (n >= N) continue;
                for (int i = 0; i < 8; i++) {
                    Unit u;
                    F::template load_unit<false>(chunk.data(), r, i, u, meta);
                    for (int e = 0; e < 32; e++) {
                        const int64_t k = kc * CHUNK_K + 32 * i + e;
                        if (k >= K) continue;
                        q[n * K + k] = (int8_t)(unit_elem(u, e) - u.off[e >> 4]);
                    }
                    for (int h = 0; h < 2; h++) {
                        const int64_t k0 = kc * CHUNK_K + 32 * i + 16 * h;
                        if (k0 >= K) continue;
                        const int64_t p = (F::SUB == 32) ? (kc * nsub + i) : (kc * nsub + 2 * i + h);
                        a[n * (K / F::SUB) + p] = u.a[h];
                        b[n * (K / F::SUB) + p] = u.b[h];
                    }
                }
            }
        }
    return 0;
}

extern "C" int hostfmt_decompose(int ggml_type, const uint8_t* blocks, int64_t N, int64_t K, int8_t* q, float* a, float* b) {
    switch (ggml_type) {
        case 12: return dump<FmtQ4K>(blocks, N, K, 1, q, a, b);
        case 14: return dump<FmtQ6K>(blocks, N, K, 1, q, a, b);
        case 8: return dump<FmtQ8_0>(blocks, N, K, 1, q, a, b);
        case 13: return dump<FmtQ5K>(blocks, N, K, 1, q, a, b);
        case 3: return dump<FmtQ4_1>(blocks, N, K, 1, q, a, b);
        case 7: return dump<FmtQ5_1>(blocks, N, K, 1, q, a, b);
        case 10: return dump<FmtQ2K>(blocks, N, K, 1, q, a, b);
        case 11: return dump<FmtQ3K>(blocks, N, K, 1, q, a, b);
        case 23: return dump<FmtIQ4XS>(blocks, N, K, 1, q, a, b);
        case 35: return dump<FmtTQ2_0>(blocks, N, K, 1, q, a, b);
        default: return -1;
    }
}

// source adaptors: repack with the adaptor, read back with the family it targets
template <class S, class F>
static int dump_adapt(const uint8_t* blocks, int64_t N, int64_t K, int gpc, int8_t* q, float* a, float* b) {
    struct Mixed : F {
        static constexpr int src_block_elems() { return S::src_block_elems(); }
        static constexpr int src_block_bytes() { return S::src_block_bytes(); }
        static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta m) { S::repack_row(src, nvalid, chunk, r, m); }
    };
    return dump<Mixed>(blocks, N, K, gpc, q, a, b);
}
extern "C" int hostfmt_decompose_adapted(int ggml_type, const uint8_t* blocks, int64_t N, int64_t K, int8_t* q, float* a, float* b) {
    switch (ggml_type) {
        case 2: return dump_adapt<SrcQ4_0, FmtG4>(blocks, N, K, 8, q, a, b);
        case 6: return dump_adapt<SrcQ5_0, FmtQ8_0>(blocks, N, K, 1, q, a, b);
        case 20: return dump_adapt<SrcIQ4NL, FmtQ8_0>(blocks, N, K, 1, q, a, b);
        case 34: return dump_adapt<SrcTQ1_0, FmtTQ2_0>(blocks, N, K, 1, q, a, b);
        case 16: return dump_adapt<SrcIQ2XXS, FmtI8S>(blocks, N, K, 3, q, a, b);
        case 17: return dump_adapt<SrcIQ2XS, FmtI8S>(blocks, N, K, 3, q, a, b);
        case 18: return dump_adapt<SrcIQ3XXS, FmtI8S>(blocks, N, K, 2, q, a, b);
        case 22: return dump_adapt<SrcIQ2S, FmtI8S>(blocks, N, K, 3, q, a, b);
        case 21: return dump_adapt<SrcIQ3S, FmtI8S>(blocks, N, K, 0, q, a, b);
        case 19: return dump_adapt<SrcIQ1S, FmtI8S>(blocks, N, K, 3, q, a, b);
        case 29: return dump_adapt<SrcIQ1M, FmtI8S>(blocks, N, K, 3, q, a, b);
        default: return -1;
    }
}
