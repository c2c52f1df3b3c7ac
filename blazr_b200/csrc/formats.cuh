// formats.cuh -- kernel-native tile layouts of the quantized weight formats.
//
// Device layout (DESIGN.md "Data layout in HBM"): a logical [N,K] weight is cut into tiles of
// TILE_ROWS=128 rows x CHUNK_K=256 k ("chunks"), stored tile-row-major then k-chunk-major, each chunk one
// contiguous, 16-byte aligned byte string that a single cp.async.bulk (TMA) copy brings into shared memory.
// Inside a chunk the canonical ggml block is split into SoA planes over the 128 rows (so 210-byte Q6_K and
// 34-byte Q8_0 blocks become 16-byte aligned without padding: device bytes == canonical bytes), the 4/6-bit
// payload is re-ordered so that one 16-byte "unit" holds 32 consecutive k of one row (low nibbles = first 16,
// high nibbles = last 16), and 16-byte units are XOR-swizzled by the row index so that BOTH access patterns
// are bank-conflict free:  (a) dp4a matvec: 8 lanes x 16 B walk one row (b) tcgen05 GEMM: thread == row.
//
// Every format exposes the same micro-interface:
//   repack_row(src canonical bytes of this row's 256 k, n_valid, chunk, r, meta)   one-off at upload
//   load_unit<SMEM>(chunk, r, i, Unit&, meta)    -> 32 integer weights (as 8 packed byte words) of unit i
//        plus the affine map  W = a[h] * (v - off[h]) - b[h]   for the two 16-element halves h.
// which is the decomposition the oracle uses (oracle/quant_oracle.c decompose_block), so dequantized weights
// and integer partials are bit-comparable.
#pragma once
#ifdef B200Q_HOST_CHECK
#include "hostcheck/host_shim.h"  // CPU build of the layout logic for tests/test_host_formats.py
#else
#include "common.cuh"
#endif

namespace b200q {

struct Unit {
    uint32_t v[8];  // v[k] byte c = integer weight of element 4k+c of the unit (k<4: first half, k>=4: second half)
    float a[2], b[2];
    int off[2];
};

struct FmtMeta {
    int gpc;  // AWQ/GPTQ: scale groups per 256-k chunk (1,2,4,8); unused otherwise
};

template <bool SMEM>
__device__ __forceinline__ uint4 ld16(const uint8_t* p) {
    if constexpr (SMEM) return lds128(p);
    else return *reinterpret_cast<const uint4*>(p);
}
template <bool SMEM>
__device__ __forceinline__ uint2 ld8(const uint8_t* p) {
    if constexpr (SMEM) return lds64(p);
    else return *reinterpret_cast<const uint2*>(p);
}
__device__ __forceinline__ uint32_t ld4(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint16_t ld2(const uint8_t* p) { return *reinterpret_cast<const uint16_t*>(p); }

__device__ __forceinline__ int swz8(int r, int u) { return u ^ (r & 7); }          // planes with >= 128 B per row
__device__ __forceinline__ int swz4(int r, int u) { return u ^ ((r >> 1) & 3); }   // planes with 64 B per row

// Q4_K / Q5_K 6-bit scale+min unpack from the canonical 12 bytes
__device__ __forceinline__ void k4_scale_min(int j, const uint8_t* s, int& sc, int& m) {
    if (j < 4) {
        sc = s[j] & 63;
        m = s[j + 4] & 63;
    } else {
        sc = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4);
        m = (s[j + 4] >> 4) | ((s[j] >> 6) << 4);
    }
}

// write 32 4-bit values q[0..31] as one 16-byte unit: byte b = q[b] | q[16+b] << 4
__device__ __forceinline__ void store_nib_unit(uint8_t* dst, const uint8_t* q) {
#pragma unroll
    for (int b = 0; b < 16; b++) dst[b] = (uint8_t)((q[b] & 0xF) | ((q[16 + b] & 0xF) << 4));
}

// ------------------------------------------------------------------------------------------------
// Q4_K : canonical [f16 d][f16 dmin][u8 s[12]][u8 qs[128]]  (144 B / 256)
// chunk: QS plane 128 rows x 128 B | HDR plane 128 rows x 16 B: [f16 d][f16 dmin][sc'|m' 12 B]
//   sc'/m' re-encoding (same 12 bytes, lane-uniform extraction): byte j = sc_j | (m_j & 3) << 6 (j<8),
//   byte 8 + j/2 nibble j%2 = m_j >> 2.
// ------------------------------------------------------------------------------------------------
struct FmtQ4K {
    static constexpr int FAMILY = 1, SUB = 32;
    static constexpr bool NIB = true;   // 4-bit payload: load_unit<.., RAWHI> available
    static constexpr bool SIGNED = false;  // unit bytes are signed int8 (else unsigned small integers)
    static constexpr bool HAS_MIN = true;
    static constexpr int QS = 0, HDR = 128 * 128;
    __host__ __device__ static constexpr int chunk_bytes(int) { return 128 * 144; }
    __host__ __device__ static constexpr int src_block_elems() { return 256; }
    __host__ __device__ static constexpr int src_block_bytes() { return 144; }

    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t* qs = chunk + QS + r * 128;
        uint8_t* hdr = chunk + HDR + r * 16;
        if (nvalid <= 0) {
            for (int j = 0; j < 128; j++) qs[j] = 0;
            for (int j = 0; j < 16; j++) hdr[j] = 0;
            return;
        }
        for (int j = 0; j < 4; j++) hdr[j] = src[j];
        const uint8_t* s = src + 4;
        const uint8_t* q = src + 16;
        int sc[8], m[8];
        for (int j = 0; j < 8; j++) k4_scale_min(j, s, sc[j], m[j]);
        for (int j = 0; j < 8; j++) hdr[4 + j] = (uint8_t)(sc[j] | ((m[j] & 3) << 6));
        for (int j = 0; j < 4; j++) hdr[12 + j] = (uint8_t)((m[2 * j] >> 2) | ((m[2 * j + 1] >> 2) << 4));
        for (int i = 0; i < 8; i++) {  // unit i == sub-block i: chunk c = i/2, high nibble iff i odd
            uint8_t v[32];
            int c = i >> 1, hi = i & 1;
            for (int l = 0; l < 32; l++) v[l] = hi ? (q[32 * c + l] >> 4) : (q[32 * c + l] & 0xF);
            store_nib_unit(qs + 16 * swz8(r, i), v);
        }
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta) {
        uint4 w = ld16<SMEM>(chunk + QS + r * 128 + 16 * swz8(r, i));
        u.v[0] = w.x & 0x0F0F0F0Fu; u.v[1] = w.y & 0x0F0F0F0Fu; u.v[2] = w.z & 0x0F0F0F0Fu; u.v[3] = w.w & 0x0F0F0F0Fu;
        if constexpr (RAWHI) {  // matvec: high nibbles stay in place (16 x value), the integer dot is shifted back once
            u.v[4] = w.x & 0xF0F0F0F0u; u.v[5] = w.y & 0xF0F0F0F0u; u.v[6] = w.z & 0xF0F0F0F0u; u.v[7] = w.w & 0xF0F0F0F0u;
        } else {
            u.v[4] = (w.x >> 4) & 0x0F0F0F0Fu; u.v[5] = (w.y >> 4) & 0x0F0F0F0Fu;
            u.v[6] = (w.z >> 4) & 0x0F0F0F0Fu; u.v[7] = (w.w >> 4) & 0x0F0F0F0Fu;
        }
        const uint8_t* hdr = chunk + HDR + r * 16;
        uint32_t dd = ld4(hdr);
        float d = half_bits_to_float((uint16_t)(dd & 0xFFFF)), dmin = half_bits_to_float((uint16_t)(dd >> 16));
        uint32_t b1 = hdr[4 + i], b2 = hdr[12 + (i >> 1)];
        int sc = b1 & 63;
        int m = (b1 >> 6) | (((b2 >> (4 * (i & 1))) & 0xF) << 2);
        u.a[0] = u.a[1] = __fmul_rn(d, (float)sc);
        u.b[0] = u.b[1] = __fmul_rn(dmin, (float)m);
        u.off[0] = u.off[1] = 0;
    }
};

// ------------------------------------------------------------------------------------------------
// Q6_K : canonical [u8 ql[128]][u8 qh[64]][i8 sc[16]][f16 d]  (210 B / 256)
// chunk: QL 128x128 B (low 4 bits, unit order) | QH 128x64 B | SC 128x16 B | D 128x2 B
//   qh' unit i = 2 words: word h serves elements 32i+16h+e (e<16): bits [8(e%4)+2(e/4) .. +1] = v>>4
// ------------------------------------------------------------------------------------------------
struct FmtQ6K {
    static constexpr int FAMILY = 2, SUB = 16;
    static constexpr bool NIB = false;
    static constexpr bool SIGNED = false;  // unit bytes are signed int8 (else unsigned small integers)
    static constexpr bool HAS_MIN = false;
    static constexpr int QL = 0, QH = 128 * 128, SC = QH + 128 * 64, D = SC + 128 * 16;
    __host__ __device__ static constexpr int chunk_bytes(int) { return 128 * 210; }
    __host__ __device__ static constexpr int src_block_elems() { return 256; }
    __host__ __device__ static constexpr int src_block_bytes() { return 210; }

    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t* ql = chunk + QL + r * 128;
        uint8_t* qh = chunk + QH + r * 64;
        uint8_t* sc = chunk + SC + r * 16;
        uint8_t* dd = chunk + D + r * 2;
        if (nvalid <= 0) {
            for (int j = 0; j < 128; j++) ql[j] = 0;
            for (int j = 0; j < 64; j++) qh[j] = 0;
            for (int j = 0; j < 16; j++) sc[j] = 0;
            dd[0] = dd[1] = 0;
            return;
        }
        const uint8_t* L0 = src; const uint8_t* H0 = src + 128;
        for (int j = 0; j < 16; j++) sc[j] = src[192 + j];
        dd[0] = src[208]; dd[1] = src[209];
        for (int i = 0; i < 8; i++) {
            uint8_t v[32];
            for (int l = 0; l < 32; l++) {
                int e = 32 * i + l, h = e >> 7, rr = (e & 127) >> 5, ll = e & 31;
                const uint8_t* L = L0 + 64 * h; const uint8_t* H = H0 + 32 * h;
                int lo = (rr & 1) ? L[ll + 32] : L[ll];
                lo = (rr & 2) ? (lo >> 4) : (lo & 15);
                int hi = (H[ll] >> (2 * rr)) & 3;
                v[l] = (uint8_t)(lo | (hi << 4));
            }
            store_nib_unit(ql + 16 * swz8(r, i), v);
            uint32_t w[2] = {0u, 0u};
            for (int h = 0; h < 2; h++)
                for (int e = 0; e < 16; e++) w[h] |= (uint32_t)(v[16 * h + e] >> 4) << (8 * (e & 3) + 2 * (e >> 2));
            uint8_t* dst = qh + 16 * swz4(r, i >> 1) + 8 * (i & 1);
            for (int h = 0; h < 2; h++)
                for (int c = 0; c < 4; c++) dst[4 * h + c] = (uint8_t)(w[h] >> (8 * c));
        }
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta) {
        uint4 w = ld16<SMEM>(chunk + QL + r * 128 + 16 * swz8(r, i));
        uint2 hq = ld8<SMEM>(chunk + QH + r * 64 + 16 * swz4(r, i >> 1) + 8 * (i & 1));
        const uint32_t M4 = 0x0F0F0F0Fu, MH = 0x30303030u;
        u.v[0] = (w.x & M4) | ((hq.x << 4) & MH);
        u.v[1] = (w.y & M4) | ((hq.x << 2) & MH);
        u.v[2] = (w.z & M4) | (hq.x & MH);
        u.v[3] = (w.w & M4) | ((hq.x >> 2) & MH);
        u.v[4] = ((w.x >> 4) & M4) | ((hq.y << 4) & MH);
        u.v[5] = ((w.y >> 4) & M4) | ((hq.y << 2) & MH);
        u.v[6] = ((w.z >> 4) & M4) | (hq.y & MH);
        u.v[7] = ((w.w >> 4) & M4) | ((hq.y >> 2) & MH);
        uint16_t s2 = ld2(chunk + SC + r * 16 + 2 * i);
        float d = half_bits_to_float(ld2(chunk + D + r * 2));
        u.a[0] = __fmul_rn(d, (float)(int8_t)(s2 & 0xFF));
        u.a[1] = __fmul_rn(d, (float)(int8_t)(s2 >> 8));
        u.b[0] = u.b[1] = 0.0f;
        u.off[0] = u.off[1] = 32;
    }
};

// ------------------------------------------------------------------------------------------------
// Q8_0 : canonical 8 x [f16 d][i8 qs[32]]  (272 B / 256)
// chunk: QS 128x256 B | D 128x16 B (8 x f16)
// ------------------------------------------------------------------------------------------------
struct FmtQ8_0 {
    static constexpr int FAMILY = 3, SUB = 32;
    static constexpr bool NIB = false;
    static constexpr bool SIGNED = true;  // unit bytes are signed int8 (else unsigned small integers)
    static constexpr bool HAS_MIN = false;
    static constexpr int QS = 0, D = 128 * 256;
    __host__ __device__ static constexpr int chunk_bytes(int) { return 128 * 272; }
    __host__ __device__ static constexpr int src_block_elems() { return 32; }
    __host__ __device__ static constexpr int src_block_bytes() { return 34; }

    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t* qs = chunk + QS + r * 256;
        uint8_t* dd = chunk + D + r * 16;
        for (int blk = 0; blk < 8; blk++) {
            const uint8_t* s = src + 34 * blk;
            bool ok = blk < nvalid;
            dd[2 * blk] = ok ? s[0] : 0; dd[2 * blk + 1] = ok ? s[1] : 0;
            for (int h = 0; h < 2; h++) {
                uint8_t* dst = qs + 16 * swz8(r, 2 * blk + h);
                for (int b = 0; b < 16; b++) dst[b] = ok ? s[2 + 16 * h + b] : 0;
            }
        }
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta) {
        uint4 w0 = ld16<SMEM>(chunk + QS + r * 256 + 16 * swz8(r, 2 * i));
        uint4 w1 = ld16<SMEM>(chunk + QS + r * 256 + 16 * swz8(r, 2 * i + 1));
        u.v[0] = w0.x; u.v[1] = w0.y; u.v[2] = w0.z; u.v[3] = w0.w;
        u.v[4] = w1.x; u.v[5] = w1.y; u.v[6] = w1.z; u.v[7] = w1.w;
        float d = half_bits_to_float(ld2(chunk + D + r * 16 + 2 * i));
        u.a[0] = u.a[1] = d;
        u.b[0] = u.b[1] = 0.0f;
        u.off[0] = u.off[1] = 0;
    }
};

// ------------------------------------------------------------------------------------------------
// G4 : AWQ / GPTQ INT4 group quantisation, W = (q - z) * s with z an integer zero point.
// chunk: QS 128x128 B (unit order) | SC 128 x gpc x f16 | Z 128 x gpc x u8       gpc = groups per 256 k
// ------------------------------------------------------------------------------------------------
struct FmtG4 {
    static constexpr int FAMILY = 4, SUB = 32;
    static constexpr bool NIB = true;
    static constexpr bool SIGNED = false;  // unit bytes are signed int8 (else unsigned small integers)
    static constexpr bool HAS_MIN = false;
    static constexpr int QS = 0, SC = 128 * 128;
    __host__ __device__ static constexpr int chunk_bytes(int gpc) { return 128 * 128 + 128 * gpc * 3; }
    __host__ __device__ static constexpr int z_off(int gpc) { return SC + 128 * gpc * 2; }

    // q[256] 4-bit values of row r in k order, sc[gpc] f16 bits, z[gpc] integer zero points
    __device__ static void store_row(const uint8_t* q, const uint16_t* sc, const uint8_t* z, uint8_t* chunk, int r, int gpc) {
        uint8_t* qs = chunk + QS + r * 128;
        for (int i = 0; i < 8; i++) store_nib_unit(qs + 16 * swz8(r, i), q + 32 * i);
        uint8_t* s = chunk + SC + r * gpc * 2;
        uint8_t* zz = chunk + z_off(gpc) + r * gpc;
        for (int g = 0; g < gpc; g++) { s[2 * g] = (uint8_t)(sc[g] & 0xFF); s[2 * g + 1] = (uint8_t)(sc[g] >> 8); zz[g] = z[g]; }
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta meta) {
        uint4 w = ld16<SMEM>(chunk + QS + r * 128 + 16 * swz8(r, i));
        u.v[0] = w.x & 0x0F0F0F0Fu; u.v[1] = w.y & 0x0F0F0F0Fu; u.v[2] = w.z & 0x0F0F0F0Fu; u.v[3] = w.w & 0x0F0F0F0Fu;
        if constexpr (RAWHI) {  // matvec: high nibbles stay in place (16 x value), the integer dot is shifted back once
            u.v[4] = w.x & 0xF0F0F0F0u; u.v[5] = w.y & 0xF0F0F0F0u; u.v[6] = w.z & 0xF0F0F0F0u; u.v[7] = w.w & 0xF0F0F0F0u;
        } else {
            u.v[4] = (w.x >> 4) & 0x0F0F0F0Fu; u.v[5] = (w.y >> 4) & 0x0F0F0F0Fu;
            u.v[6] = (w.z >> 4) & 0x0F0F0F0Fu; u.v[7] = (w.w >> 4) & 0x0F0F0F0Fu;
        }
        int gq = (i * meta.gpc) >> 3;
        float s = half_bits_to_float(ld2(chunk + SC + (r * meta.gpc + gq) * 2));
        int z = chunk[z_off(meta.gpc) + r * meta.gpc + gq];
        u.a[0] = u.a[1] = s;
        u.b[0] = u.b[1] = 0.0f;
        u.off[0] = u.off[1] = z;
    }
};

// ------------------------------------------------------------------------------------------------
// Formats added after the headline three ("the remaining K formats", BASELINE north_star).  Same micro-interface, so
// every kernel (dequantize, integer partials, matvec, grouped matvec, tcgen05 GEMM) is instantiated unchanged.  Their
// repack_row / load_unit are verified on the CPU against the oracle (hostcheck/, tests/test_host_formats.py).
// High-bit planes are re-encoded so that word K of a unit gets its 4 bits with one shift + mask:
//   hb' bit (8 c + K) = extra bit of element 4 K + c         (K = 0..7 unit words, c = 0..3 bytes)
// and 2-bit payloads so that   v[4 h + kk] = (W[h] >> 2 kk) & 0x03030303   (h = 16-element half, kk = 0..3).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int swzw(int r, int i) { return i ^ ((r >> 2) & 7); }   // planes with one 4-byte word per unit (32 B rows)
__device__ __forceinline__ int swzd(int r, int i) { return i ^ ((r >> 1) & 7); }   // planes with 8 bytes per unit (64 B rows)

__device__ __forceinline__ uint32_t pack_hibits(const uint8_t* bit /*[32] 0/1*/) {
    uint32_t w = 0;
    for (int K = 0; K < 8; K++)
        for (int c = 0; c < 4; c++) w |= (uint32_t)(bit[4 * K + c] & 1) << (8 * c + K);
    return w;
}
__device__ __forceinline__ void store_u32(uint8_t* dst, uint32_t w) {
    for (int c = 0; c < 4; c++) dst[c] = (uint8_t)(w >> (8 * c));
}
// 32 two-bit values -> 2 words (see above)
__device__ __forceinline__ void store_2bit_unit(uint8_t* dst, const uint8_t* q /*[32]*/) {
    for (int h = 0; h < 2; h++) {
        uint32_t w = 0;
        for (int kk = 0; kk < 4; kk++)
            for (int c = 0; c < 4; c++) w |= (uint32_t)(q[16 * h + 4 * kk + c] & 3) << (8 * c + 2 * kk);
        store_u32(dst + 4 * h, w);
    }
}

// Q5_K : canonical [f16 d][f16 dmin][u8 s[12]][u8 qh[32]][u8 qs[128]]  (176 B / 256)
// chunk: QS 128x128 B (low nibbles, unit order) | QH 128x32 B (hb' words) | HDR 128x16 B (as Q4_K)
struct FmtQ5K {
    static constexpr int FAMILY = 5, SUB = 32;
    static constexpr bool NIB = false;
    static constexpr bool SIGNED = false;
    static constexpr bool HAS_MIN = true;
    static constexpr int QS = 0, QH = 128 * 128, HDR = QH + 128 * 32;
    __host__ __device__ static constexpr int chunk_bytes(int) { return 128 * 176; }
    __host__ __device__ static constexpr int src_block_elems() { return 256; }
    __host__ __device__ static constexpr int src_block_bytes() { return 176; }

    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t* qs = chunk + QS + r * 128;
        uint8_t* qh = chunk + QH + r * 32;
        uint8_t* hdr = chunk + HDR + r * 16;
        if (nvalid <= 0) {
            for (int j = 0; j < 128; j++) qs[j] = 0;
            for (int j = 0; j < 32; j++) qh[j] = 0;
            for (int j = 0; j < 16; j++) hdr[j] = 0;
            return;
        }
        for (int j = 0; j < 4; j++) hdr[j] = src[j];
        const uint8_t* s = src + 4;
        const uint8_t* H = src + 16;
        const uint8_t* q = src + 48;
        int sc[8], m[8];
        for (int j = 0; j < 8; j++) k4_scale_min(j, s, sc[j], m[j]);
        for (int j = 0; j < 8; j++) hdr[4 + j] = (uint8_t)(sc[j] | ((m[j] & 3) << 6));
        for (int j = 0; j < 4; j++) hdr[12 + j] = (uint8_t)((m[2 * j] >> 2) | ((m[2 * j + 1] >> 2) << 4));
        for (int i = 0; i < 8; i++) {
            uint8_t v[32], hb[32];
            const int c = i >> 1, hi = i & 1;
            for (int l = 0; l < 32; l++) {
                v[l] = hi ? (q[32 * c + l] >> 4) : (q[32 * c + l] & 0xF);
                hb[l] = (H[l] >> (2 * c + hi)) & 1;
            }
            store_nib_unit(qs + 16 * swz8(r, i), v);
            store_u32(qh + 4 * swzw(r, i), pack_hibits(hb));
        }
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta) {
        uint4 w = ld16<SMEM>(chunk + QS + r * 128 + 16 * swz8(r, i));
        const uint32_t hq = ld4(chunk + QH + r * 32 + 4 * swzw(r, i));
        const uint32_t M4 = 0x0F0F0F0Fu, M1 = 0x01010101u;
        u.v[0] = (w.x & M4) | (((hq >> 0) & M1) << 4); u.v[1] = (w.y & M4) | (((hq >> 1) & M1) << 4);
        u.v[2] = (w.z & M4) | (((hq >> 2) & M1) << 4); u.v[3] = (w.w & M4) | (((hq >> 3) & M1) << 4);
        u.v[4] = ((w.x >> 4) & M4) | (((hq >> 4) & M1) << 4); u.v[5] = ((w.y >> 4) & M4) | (((hq >> 5) & M1) << 4);
        u.v[6] = ((w.z >> 4) & M4) | (((hq >> 6) & M1) << 4); u.v[7] = ((w.w >> 4) & M4) | (((hq >> 7) & M1) << 4);
        const uint8_t* hdr = chunk + HDR + r * 16;
        uint32_t dd = ld4(hdr);
        float d = half_bits_to_float((uint16_t)(dd & 0xFFFF)), dmin = half_bits_to_float((uint16_t)(dd >> 16));
        uint32_t b1 = hdr[4 + i], b2 = hdr[12 + (i >> 1)];
        int sc = b1 & 63;
        int m = (b1 >> 6) | (((b2 >> (4 * (i & 1))) & 0xF) << 2);
        u.a[0] = u.a[1] = __fmul_rn(d, (float)sc);
        u.b[0] = u.b[1] = __fmul_rn(dmin, (float)m);
        u.off[0] = u.off[1] = 0;
    }
};

// Q4_1 : canonical 8 x [f16 d][f16 m][u8 qs[16]]  (160 B / 256),  w = d q + m
// chunk: QS 128x128 B (block i's 16 bytes as stored: byte b = q[b] | q[16+b] << 4) | DM 128x32 B (8 x (d, m))
struct FmtQ4_1 {
    static constexpr int FAMILY = 7, SUB = 32;
    static constexpr bool NIB = true;
    static constexpr bool SIGNED = false;
    static constexpr bool HAS_MIN = true;
    static constexpr int QS = 0, DM = 128 * 128;
    __host__ __device__ static constexpr int chunk_bytes(int) { return 128 * 160; }
    __host__ __device__ static constexpr int src_block_elems() { return 32; }
    __host__ __device__ static constexpr int src_block_bytes() { return 20; }

    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t* qs = chunk + QS + r * 128;
        uint8_t* dm = chunk + DM + r * 32;
        for (int blk = 0; blk < 8; blk++) {
            const uint8_t* s = src + 20 * blk;
            const bool ok = blk < nvalid;
            uint8_t* d4 = dm + 4 * swzw(r, blk);
            for (int c = 0; c < 4; c++) d4[c] = ok ? s[c] : 0;
            uint8_t* dst = qs + 16 * swz8(r, blk);
            for (int b = 0; b < 16; b++) dst[b] = ok ? s[4 + b] : 0;
        }
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta) {
        uint4 w = ld16<SMEM>(chunk + QS + r * 128 + 16 * swz8(r, i));
        u.v[0] = w.x & 0x0F0F0F0Fu; u.v[1] = w.y & 0x0F0F0F0Fu; u.v[2] = w.z & 0x0F0F0F0Fu; u.v[3] = w.w & 0x0F0F0F0Fu;
        if constexpr (RAWHI) {
            u.v[4] = w.x & 0xF0F0F0F0u; u.v[5] = w.y & 0xF0F0F0F0u; u.v[6] = w.z & 0xF0F0F0F0u; u.v[7] = w.w & 0xF0F0F0F0u;
        } else {
            u.v[4] = (w.x >> 4) & 0x0F0F0F0Fu; u.v[5] = (w.y >> 4) & 0x0F0F0F0Fu;
            u.v[6] = (w.z >> 4) & 0x0F0F0F0Fu; u.v[7] = (w.w >> 4) & 0x0F0F0F0Fu;
        }
        const uint32_t dm = ld4(chunk + DM + r * 32 + 4 * swzw(r, i));
        u.a[0] = u.a[1] = half_bits_to_float((uint16_t)(dm & 0xFFFF));
        u.b[0] = u.b[1] = -half_bits_to_float((uint16_t)(dm >> 16));
        u.off[0] = u.off[1] = 0;
    }
};

// Q5_1 : canonical 8 x [f16 d][f16 m][u32 qh][u8 qs[16]]  (192 B / 256),  w = d q5 + m
// chunk: QS 128x128 B | QH 128x32 B (hb' words) | DM 128x32 B
struct FmtQ5_1 {
    static constexpr int FAMILY = 9, SUB = 32;
    static constexpr bool NIB = false;
    static constexpr bool SIGNED = false;
    static constexpr bool HAS_MIN = true;
    static constexpr int QS = 0, QH = 128 * 128, DM = QH + 128 * 32;
    __host__ __device__ static constexpr int chunk_bytes(int) { return 128 * 192; }
    __host__ __device__ static constexpr int src_block_elems() { return 32; }
    __host__ __device__ static constexpr int src_block_bytes() { return 24; }

    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t* qs = chunk + QS + r * 128;
        uint8_t* qh = chunk + QH + r * 32;
        uint8_t* dm = chunk + DM + r * 32;
        for (int blk = 0; blk < 8; blk++) {
            const uint8_t* s = src + 24 * blk;
            const bool ok = blk < nvalid;
            uint8_t* d4 = dm + 4 * swzw(r, blk);
            for (int c = 0; c < 4; c++) d4[c] = ok ? s[c] : 0;
            const uint32_t h = ok ? ((uint32_t)s[4] | ((uint32_t)s[5] << 8) | ((uint32_t)s[6] << 16) | ((uint32_t)s[7] << 24)) : 0u;
            uint8_t hb[32];
            for (int e = 0; e < 32; e++) hb[e] = (h >> e) & 1;   // canonical: bit e of qh = 5th bit of element e
            store_u32(qh + 4 * swzw(r, blk), pack_hibits(hb));
            uint8_t* dst = qs + 16 * swz8(r, blk);
            for (int b = 0; b < 16; b++) dst[b] = ok ? s[8 + b] : 0;
        }
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta) {
        uint4 w = ld16<SMEM>(chunk + QS + r * 128 + 16 * swz8(r, i));
        const uint32_t hq = ld4(chunk + QH + r * 32 + 4 * swzw(r, i));
        const uint32_t M4 = 0x0F0F0F0Fu, M1 = 0x01010101u;
        u.v[0] = (w.x & M4) | (((hq >> 0) & M1) << 4); u.v[1] = (w.y & M4) | (((hq >> 1) & M1) << 4);
        u.v[2] = (w.z & M4) | (((hq >> 2) & M1) << 4); u.v[3] = (w.w & M4) | (((hq >> 3) & M1) << 4);
        u.v[4] = ((w.x >> 4) & M4) | (((hq >> 4) & M1) << 4); u.v[5] = ((w.y >> 4) & M4) | (((hq >> 5) & M1) << 4);
        u.v[6] = ((w.z >> 4) & M4) | (((hq >> 6) & M1) << 4); u.v[7] = ((w.w >> 4) & M4) | (((hq >> 7) & M1) << 4);
        const uint32_t dm = ld4(chunk + DM + r * 32 + 4 * swzw(r, i));
        u.a[0] = u.a[1] = half_bits_to_float((uint16_t)(dm & 0xFFFF));
        u.b[0] = u.b[1] = -half_bits_to_float((uint16_t)(dm >> 16));
        u.off[0] = u.off[1] = 0;
    }
};

// Q2_K : canonical [u8 scales[16] (4-bit sc | 4-bit m << 4)][u8 qs[64]][f16 d][f16 dmin]  (84 B / 256), 16 sub-blocks of 16
// chunk: Q2 128x64 B (unit = 2 words) | SC 128x16 B (as stored) | DD 128x4 B
struct FmtQ2K {
    static constexpr int FAMILY = 10, SUB = 16;
    static constexpr bool NIB = false;
    static constexpr bool SIGNED = false;
    static constexpr bool HAS_MIN = true;
    static constexpr int Q2 = 0, SC = 128 * 64, DD = SC + 128 * 16;
    __host__ __device__ static constexpr int chunk_bytes(int) { return 128 * 84; }
    __host__ __device__ static constexpr int src_block_elems() { return 256; }
    __host__ __device__ static constexpr int src_block_bytes() { return 84; }

    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t* q2 = chunk + Q2 + r * 64;
        uint8_t* sc = chunk + SC + r * 16;
        uint8_t* dd = chunk + DD + r * 4;
        if (nvalid <= 0) {
            for (int j = 0; j < 64; j++) q2[j] = 0;
            for (int j = 0; j < 16; j++) sc[j] = 0;
            for (int j = 0; j < 4; j++) dd[j] = 0;
            return;
        }
        for (int j = 0; j < 16; j++) sc[j] = src[j];
        for (int j = 0; j < 4; j++) dd[j] = src[80 + j];
        const uint8_t* qs = src + 16;
        for (int i = 0; i < 8; i++) {
            uint8_t v[32];
            for (int l = 0; l < 32; l++) {
                const int e = 32 * i + l, n = e >> 7, j = (e & 127) >> 5, ll = e & 31;
                v[l] = (qs[32 * n + ll] >> (2 * j)) & 3;
            }
            store_2bit_unit(q2 + 8 * swzd(r, i), v);
        }
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta) {
        const uint2 q = ld8<SMEM>(chunk + Q2 + r * 64 + 8 * swzd(r, i));
        const uint32_t M2 = 0x03030303u;
        u.v[0] = q.x & M2; u.v[1] = (q.x >> 2) & M2; u.v[2] = (q.x >> 4) & M2; u.v[3] = (q.x >> 6) & M2;
        u.v[4] = q.y & M2; u.v[5] = (q.y >> 2) & M2; u.v[6] = (q.y >> 4) & M2; u.v[7] = (q.y >> 6) & M2;
        const uint16_t s2 = ld2(chunk + SC + r * 16 + 2 * i);
        const uint32_t dd = ld4(chunk + DD + r * 4);
        const float d = half_bits_to_float((uint16_t)(dd & 0xFFFF)), dmin = half_bits_to_float((uint16_t)(dd >> 16));
        u.a[0] = __fmul_rn(d, (float)(s2 & 0xF));
        u.b[0] = __fmul_rn(dmin, (float)((s2 >> 4) & 0xF));
        u.a[1] = __fmul_rn(d, (float)((s2 >> 8) & 0xF));
        u.b[1] = __fmul_rn(dmin, (float)(s2 >> 12));
        u.off[0] = u.off[1] = 0;
    }
};

// Q3_K : canonical [u8 hmask[32]][u8 qs[64]][u8 scales[12]][f16 d]  (110 B / 256), 16 sub-blocks of 16,
//        w = d (sc - 32) (q3 - 4),  q3 = 2 low bits | hmask bit << 2
// chunk: Q2 128x64 B | HM 128x32 B (hb' words) | SC 128x12 B re-encoded: byte i = lo4(sc[2i]) | lo4(sc[2i+1]) << 4,
//        byte 8 + i/2 nibble i%2 = hi2(sc[2i]) | hi2(sc[2i+1]) << 2 | D 128x2 B
struct FmtQ3K {
    static constexpr int FAMILY = 11, SUB = 16;
    static constexpr bool NIB = false;
    static constexpr bool SIGNED = false;
    static constexpr bool HAS_MIN = false;
    static constexpr int Q2 = 0, HM = 128 * 64, SC = HM + 128 * 32, D = SC + 128 * 12;
    __host__ __device__ static constexpr int chunk_bytes(int) { return 128 * 110; }
    __host__ __device__ static constexpr int src_block_elems() { return 256; }
    __host__ __device__ static constexpr int src_block_bytes() { return 110; }

    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t* q2 = chunk + Q2 + r * 64;
        uint8_t* hm = chunk + HM + r * 32;
        uint8_t* sc = chunk + SC + r * 12;
        uint8_t* dd = chunk + D + r * 2;
        if (nvalid <= 0) {
            for (int j = 0; j < 64; j++) q2[j] = 0;
            for (int j = 0; j < 32; j++) hm[j] = 0;
            for (int j = 0; j < 12; j++) sc[j] = 0;
            dd[0] = dd[1] = 0;
            return;
        }
        const uint8_t* H = src;
        const uint8_t* qs = src + 32;
        const uint8_t* s = src + 96;
        dd[0] = src[108]; dd[1] = src[109];
        int scl[16];
        for (int i = 0; i < 16; i++) {
            const int lo = (i < 8) ? (s[i] & 0xF) : (s[i - 8] >> 4);
            const int hi = (s[8 + (i % 4)] >> (2 * (i / 4))) & 3;
            scl[i] = lo | (hi << 4);
        }
        for (int i = 0; i < 8; i++) sc[i] = (uint8_t)((scl[2 * i] & 15) | ((scl[2 * i + 1] & 15) << 4));
        for (int j = 0; j < 4; j++) {
            const int i0 = 2 * j, i1 = 2 * j + 1;
            const int n0 = (scl[2 * i0] >> 4) | ((scl[2 * i0 + 1] >> 4) << 2);
            const int n1 = (scl[2 * i1] >> 4) | ((scl[2 * i1 + 1] >> 4) << 2);
            sc[8 + j] = (uint8_t)(n0 | (n1 << 4));
        }
        for (int i = 0; i < 8; i++) {
            uint8_t v[32], hb[32];
            for (int l = 0; l < 32; l++) {
                const int e = 32 * i + l, n = e >> 7, j = (e & 127) >> 5, ll = e & 31;
                v[l] = (qs[32 * n + ll] >> (2 * j)) & 3;
                hb[l] = (H[ll] >> (4 * n + j)) & 1;
            }
            store_2bit_unit(q2 + 8 * swzd(r, i), v);
            store_u32(hm + 4 * swzw(r, i), pack_hibits(hb));
        }
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta) {
        const uint2 q = ld8<SMEM>(chunk + Q2 + r * 64 + 8 * swzd(r, i));
        const uint32_t hq = ld4(chunk + HM + r * 32 + 4 * swzw(r, i));
        const uint32_t M2 = 0x03030303u, M1 = 0x01010101u;
        u.v[0] = (q.x & M2) | (((hq >> 0) & M1) << 2); u.v[1] = ((q.x >> 2) & M2) | (((hq >> 1) & M1) << 2);
        u.v[2] = ((q.x >> 4) & M2) | (((hq >> 2) & M1) << 2); u.v[3] = ((q.x >> 6) & M2) | (((hq >> 3) & M1) << 2);
        u.v[4] = (q.y & M2) | (((hq >> 4) & M1) << 2); u.v[5] = ((q.y >> 2) & M2) | (((hq >> 5) & M1) << 2);
        u.v[6] = ((q.y >> 4) & M2) | (((hq >> 6) & M1) << 2); u.v[7] = ((q.y >> 6) & M2) | (((hq >> 7) & M1) << 2);
        const uint8_t* sc = chunk + SC + r * 12;
        const uint32_t b1 = sc[i], b2 = (uint32_t)sc[8 + (i >> 1)] >> (4 * (i & 1));
        const int s0 = (int)((b1 & 15) | ((b2 & 3) << 4)) - 32;
        const int s1 = (int)((b1 >> 4) | (((b2 >> 2) & 3) << 4)) - 32;
        const float d = half_bits_to_float(ld2(chunk + D + r * 2));
        u.a[0] = __fmul_rn(d, (float)s0);
        u.a[1] = __fmul_rn(d, (float)s1);
        u.b[0] = u.b[1] = 0.0f;
        u.off[0] = u.off[1] = 4;
    }
};

// IQ4_XS : canonical [f16 d][u16 scales_h][u8 scales_l[4]][u8 qs[128]]  (136 B / 256), 8 sub-blocks of 32,
//          w = d (ls - 32) kvalues[q]  with the 16-entry non-linear int8 codebook of IQ4_NL
// chunk: QS 128x128 B (block ib's 16 bytes as stored) | HDR 128x8 B (as stored)
// four 4-bit codes (one per byte of x, low nibble) -> four codebook bytes: two PRMTs over the table halves + select
__device__ __forceinline__ uint32_t iq4_codebook4(uint32_t x) {
    // kvalues = {-127,-104,-83,-65,-49,-35,-22,-10, 1,13,25,38,53,69,89,113}
    const uint32_t T0 = 0xBFAD9881u, T1 = 0xF6EADDCFu, T2 = 0x26190D01u, T3 = 0x71594535u;
    const uint32_t sel = (x & 7u) | ((x >> 4) & 0x70u) | ((x >> 8) & 0x700u) | ((x >> 12) & 0x7000u);
    const uint32_t lo = __byte_perm(T0, T1, sel), hi = __byte_perm(T2, T3, sel);
    const uint32_t m = ((x >> 3) & 0x01010101u) * 0xFFu;  // 0xFF in every byte whose code is >= 8
    return (hi & m) | (lo & ~m);
}
struct FmtIQ4XS {
    static constexpr int FAMILY = 13, SUB = 32;
    static constexpr bool NIB = false;
    static constexpr bool SIGNED = true;   // unit bytes are signed int8 codebook values
    static constexpr bool HAS_MIN = false;
    static constexpr int QS = 0, HDR = 128 * 128;
    __host__ __device__ static constexpr int chunk_bytes(int) { return 128 * 136; }
    __host__ __device__ static constexpr int src_block_elems() { return 256; }
    __host__ __device__ static constexpr int src_block_bytes() { return 136; }

    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t* qs = chunk + QS + r * 128;
        uint8_t* hdr = chunk + HDR + r * 8;
        if (nvalid <= 0) {
            for (int j = 0; j < 128; j++) qs[j] = 0;
            for (int j = 0; j < 8; j++) hdr[j] = 0;
            // codes 0 decode to -127: give the padding rows a zero scale (d = 0) so they contribute nothing
            return;
        }
        for (int j = 0; j < 8; j++) hdr[j] = src[j];
        for (int ib = 0; ib < 8; ib++) {
            uint8_t* dst = qs + 16 * swz8(r, ib);
            for (int b = 0; b < 16; b++) dst[b] = src[8 + 16 * ib + b];
        }
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta) {
        uint4 w = ld16<SMEM>(chunk + QS + r * 128 + 16 * swz8(r, i));
        const uint32_t M4 = 0x0F0F0F0Fu;
        u.v[0] = iq4_codebook4(w.x & M4); u.v[1] = iq4_codebook4(w.y & M4);
        u.v[2] = iq4_codebook4(w.z & M4); u.v[3] = iq4_codebook4(w.w & M4);
        u.v[4] = iq4_codebook4((w.x >> 4) & M4); u.v[5] = iq4_codebook4((w.y >> 4) & M4);
        u.v[6] = iq4_codebook4((w.z >> 4) & M4); u.v[7] = iq4_codebook4((w.w >> 4) & M4);
        const uint2 h = ld8<SMEM>(chunk + HDR + r * 8);
        const float d = half_bits_to_float((uint16_t)(h.x & 0xFFFF));
        const uint32_t sh = h.x >> 16, sl = h.y;
        const int ls = (int)(((sl >> (4 * i)) & 0xFu) | (((sh >> (2 * i)) & 3u) << 4)) - 32;
        u.a[0] = u.a[1] = __fmul_rn(d, (float)ls);
        u.b[0] = u.b[1] = 0.0f;
        u.off[0] = u.off[1] = 0;
    }
};

// TQ2_0 : ternary, canonical [u8 qs[64]][f16 d]  (66 B / 256),  w = d (q - 1),  q = 2-bit code (same element order as Q2_K)
// chunk: Q2 128x64 B (unit = 2 words) | D 128x2 B
struct FmtTQ2_0 {
    static constexpr int FAMILY = 14, SUB = 32;
    static constexpr bool NIB = false;
    static constexpr bool SIGNED = false;
    static constexpr bool HAS_MIN = false;
    static constexpr int Q2 = 0, D = 128 * 64;
    __host__ __device__ static constexpr int chunk_bytes(int) { return 128 * 66; }
    __host__ __device__ static constexpr int src_block_elems() { return 256; }
    __host__ __device__ static constexpr int src_block_bytes() { return 66; }

    // codes[256] (0..3) of one row-chunk + f16 bits of d -> planes
    __device__ static void store_row(const uint8_t* codes, uint16_t dbits, uint8_t* chunk, int r) {
        uint8_t* q2 = chunk + Q2 + r * 64;
        for (int i = 0; i < 8; i++) store_2bit_unit(q2 + 8 * swzd(r, i), codes + 32 * i);
        chunk[D + r * 2] = (uint8_t)(dbits & 0xFF);
        chunk[D + r * 2 + 1] = (uint8_t)(dbits >> 8);
    }
    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t codes[256];
        if (nvalid <= 0) {
            for (int e = 0; e < 256; e++) codes[e] = 1;   // code 1 = value 0 (and d = 0)
            store_row(codes, 0, chunk, r);
            return;
        }
        for (int e = 0; e < 256; e++) {
            const int n = e >> 7, l = (e & 127) >> 5, m = e & 31;
            codes[e] = (src[32 * n + m] >> (2 * l)) & 3;
        }
        store_row(codes, (uint16_t)(src[64] | (src[65] << 8)), chunk, r);
    }
    template <bool SMEM, bool RAWHI = false>
    __device__ __forceinline__ static void load_unit(const uint8_t* chunk, int r, int i, Unit& u, FmtMeta) {
        const uint2 q = ld8<SMEM>(chunk + Q2 + r * 64 + 8 * swzd(r, i));
        const uint32_t M2 = 0x03030303u;
        u.v[0] = q.x & M2; u.v[1] = (q.x >> 2) & M2; u.v[2] = (q.x >> 4) & M2; u.v[3] = (q.x >> 6) & M2;
        u.v[4] = q.y & M2; u.v[5] = (q.y >> 2) & M2; u.v[6] = (q.y >> 4) & M2; u.v[7] = (q.y >> 6) & M2;
        u.a[0] = u.a[1] = half_bits_to_float(ld2(chunk + D + r * 2));
        u.b[0] = u.b[1] = 0.0f;
        u.off[0] = u.off[1] = 1;
    }
};

// TQ1_0 : ternary, base-3 packed, canonical [u8 qs[48]][u8 qh[4]][f16 d]  (54 B / 256): re-encoded at upload into the
// TQ2_0 layout (2-bit codes, 66 B / 256 on the device); trit n of byte x = ((uint8)(x * 3^n) * 3) >> 8
struct SrcTQ1_0 {
    __host__ __device__ static constexpr int chunk_bytes(int gpc) { return FmtTQ2_0::chunk_bytes(gpc); }
    __host__ __device__ static constexpr int src_block_elems() { return 256; }
    __host__ __device__ static constexpr int src_block_bytes() { return 54; }
    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        uint8_t codes[256];
        if (nvalid <= 0) {
            for (int e = 0; e < 256; e++) codes[e] = 1;
            FmtTQ2_0::store_row(codes, 0, chunk, r);
            return;
        }
        const uint8_t pow3[5] = {1, 3, 9, 27, 81};
        for (int n = 0; n < 5; n++)
            for (int m = 0; m < 32; m++) codes[32 * n + m] = (uint8_t)(((uint32_t)(uint8_t)(src[m] * pow3[n]) * 3u) >> 8);
        for (int n = 0; n < 5; n++)
            for (int m = 0; m < 16; m++) codes[160 + 16 * n + m] = (uint8_t)(((uint32_t)(uint8_t)(src[32 + m] * pow3[n]) * 3u) >> 8);
        for (int n = 0; n < 4; n++)
            for (int m = 0; m < 4; m++) codes[240 + 4 * n + m] = (uint8_t)(((uint32_t)(uint8_t)(src[48 + m] * pow3[n]) * 3u) >> 8);
        FmtTQ2_0::store_row(codes, (uint16_t)(src[52] | (src[53] << 8)), chunk, r);
    }
};

// ------------------------------------------------------------------------------------------------
// Source adaptors: 32-element ggml block formats whose arithmetic is exactly expressible in an existing family are
// re-encoded at upload and then run that family's kernels (no new compute code, dequantized weights and integer
// partials stay bit-exact):
//   Q4_0   w = d (q - 8)          -> G4 with 32-wide groups: scale d, integer zero point 8          (4.75 bit/weight on device)
//   Q5_0   w = d (q5 - 16)        -> Q8_0 family: int8 = q5 - 16, same d                            (8.5 bit/weight on device)
//   IQ4_NL w = d kvalues[q]       -> Q8_0 family: int8 = kvalues[q] (non-linear 4-bit codebook)     (8.5 bit/weight on device)
// Only repack_row differs; chunk_bytes forwards to the family.
// ------------------------------------------------------------------------------------------------
struct SrcQ4_0 {
    __host__ __device__ static constexpr int chunk_bytes(int gpc) { return FmtG4::chunk_bytes(gpc); }
    __host__ __device__ static constexpr int src_block_elems() { return 32; }
    __host__ __device__ static constexpr int src_block_bytes() { return 18; }
    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta meta) {
        uint8_t q[256];
        uint16_t sc[8];
        uint8_t z[8];
        for (int blk = 0; blk < 8; blk++) {
            const uint8_t* s = src + 18 * blk;
            const bool ok = blk < nvalid;
            sc[blk] = ok ? (uint16_t)(s[0] | (s[1] << 8)) : (uint16_t)0;
            z[blk] = ok ? 8 : 0;
            for (int j = 0; j < 16; j++) {
                q[32 * blk + j] = ok ? (s[2 + j] & 0xF) : 0;
                q[32 * blk + 16 + j] = ok ? (s[2 + j] >> 4) : 0;
            }
        }
        FmtG4::store_row(q, sc, z, chunk, r, meta.gpc);
    }
};

__device__ __forceinline__ void store_q8_family_row(const int8_t (*vals)[32], const uint16_t* d, uint8_t* chunk, int r) {
    uint8_t* qs = chunk + FmtQ8_0::QS + r * 256;
    uint8_t* dd = chunk + FmtQ8_0::D + r * 16;
    for (int blk = 0; blk < 8; blk++) {
        dd[2 * blk] = (uint8_t)(d[blk] & 0xFF);
        dd[2 * blk + 1] = (uint8_t)(d[blk] >> 8);
        for (int h = 0; h < 2; h++) {
            uint8_t* dst = qs + 16 * swz8(r, 2 * blk + h);
            for (int b = 0; b < 16; b++) dst[b] = (uint8_t)vals[blk][16 * h + b];
        }
    }
}

struct SrcQ5_0 {
    __host__ __device__ static constexpr int chunk_bytes(int gpc) { return FmtQ8_0::chunk_bytes(gpc); }
    __host__ __device__ static constexpr int src_block_elems() { return 32; }
    __host__ __device__ static constexpr int src_block_bytes() { return 22; }
    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        int8_t v[8][32];
        uint16_t d[8];
        for (int blk = 0; blk < 8; blk++) {
            const uint8_t* s = src + 22 * blk;
            const bool ok = blk < nvalid;
            d[blk] = ok ? (uint16_t)(s[0] | (s[1] << 8)) : (uint16_t)0;
            const uint32_t qh = ok ? ((uint32_t)s[2] | ((uint32_t)s[3] << 8) | ((uint32_t)s[4] << 16) | ((uint32_t)s[5] << 24)) : 0u;
            for (int j = 0; j < 16; j++) {
                v[blk][j] = ok ? (int8_t)(((s[6 + j] & 0xF) | (((qh >> j) & 1) << 4)) - 16) : (int8_t)0;
                v[blk][j + 16] = ok ? (int8_t)(((s[6 + j] >> 4) | (((qh >> (j + 16)) & 1) << 4)) - 16) : (int8_t)0;
            }
        }
        store_q8_family_row(v, d, chunk, r);
    }
};

struct SrcIQ4NL {
    __host__ __device__ static constexpr int chunk_bytes(int gpc) { return FmtQ8_0::chunk_bytes(gpc); }
    __host__ __device__ static constexpr int src_block_elems() { return 32; }
    __host__ __device__ static constexpr int src_block_bytes() { return 18; }
    __device__ static void repack_row(const uint8_t* src, int nvalid, uint8_t* chunk, int r, FmtMeta) {
        const int8_t kv[16] = {-127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113};  // public ggml codebook
        int8_t v[8][32];
        uint16_t d[8];
        for (int blk = 0; blk < 8; blk++) {
            const uint8_t* s = src + 18 * blk;
            const bool ok = blk < nvalid;
            d[blk] = ok ? (uint16_t)(s[0] | (s[1] << 8)) : (uint16_t)0;
            for (int j = 0; j < 16; j++) {
                v[blk][j] = ok ? kv[s[2 + j] & 0xF] : (int8_t)0;
                v[blk][j + 16] = ok ? kv[s[2 + j] >> 4] : (int8_t)0;
            }
        }
        store_q8_family_row(v, d, chunk, r);
    }
};

// signed byte e (0..31) of a unit
__device__ __forceinline__ int unit_elem(const Unit& u, int e) { return (int)(int8_t)((u.v[e >> 2] >> (8 * (e & 3))) & 0xFF); }

}  // namespace b200q
