// host_shim.h -- lets formats.cuh (the device tile layouts: repack_row / load_unit of every format) compile for the
// HOST, so the layout logic is checked against the CPU oracle by the `-m "not gpu"` tests (tests/test_host_formats.py)
// before it ever runs on a GPU.  Test infrastructure only; nothing in the product links it.
#pragma once
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __forceinline__ inline

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };

namespace b200q {

constexpr int TILE_ROWS = 128;
constexpr int CHUNK_K = 256;

inline uint4 lds128(const void* p) { uint4 v; memcpy(&v, p, 16); return v; }
inline uint2 lds64(const void* p) { uint2 v; memcpy(&v, p, 8); return v; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }  // one IEEE rounding, never contracted

// PRMT semantics (selector nibble: bits 0-2 = byte of the 8-byte pool {x, y}, bit 3 = replicate that byte's sign)
inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t pool = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t n = (s >> (4 * i)) & 0xF;
        uint32_t b = (uint32_t)(pool >> (8 * (n & 7))) & 0xFF;
        if (n & 8) b = (b & 0x80) ? 0xFF : 0x00;
        r |= b << (8 * i);
    }
    return r;
}

// exact f16 -> f32 (subnormals, inf, nan included)
inline float half_bits_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

}  // namespace b200q
