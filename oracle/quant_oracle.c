/*
 * quant_oracle.c -- CPU ORACLE for the quantized linear-layer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under blazr_b200/ may link, import or call this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and there only as the checker / reported baseline -- never as the product path.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the third-party crates boostr 0.1.0 /
 * numr 0.5.0 (Cargo.lock:352-353, 1940-1941 of the reference), whose sources are absent from
 * /root/reference, and the reference holds no numeric test or golden vector for it.  This file
 * therefore restates
 *   (1) the public ggml block formats that blazr uploads as raw bytes
 *       (reference src/loader/gguf.rs:29-43: VarMap::from_gguf keeps ggml blocks, f32 activations
 *       src/loader/gguf.rs:305), cross-checked bit-for-bit against the independent numpy
 *       implementation gguf 0.19.0 `gguf.quants.dequantize` (fixtures under tests/golden), and
 *   (2) the AWQ / GPTQ layouts exactly as blazr's loaders hand them to the operator
 *       (reference src/loader/safetensors/awq.rs:29-32,190-226,242-263 and
 *        src/loader/safetensors/gptq.rs:1-11,198-259).
 *
 * Every weight format is decomposed as  W[k] = a_p * qi[k] - b_p  for k in sub-block p (16 or 32
 * wide), with qi a small signed integer.  That single decomposition yields
 *   - the dequantized weight (bit-exact contract #1: separate f32 multiply then subtract, no FMA;
 *     compile with -ffp-contract=off),
 *   - the integer dot partials  sum_k qi[k]*xq[k]  (bit-exact contract #2, order independent),
 *   - flavour A  (f32 activations x dequantized weights, double accumulation) and
 *     flavour B  (int8 activations, per-32 scale, integer dot + scales) results.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "iq_grids.h"  /* generated: public ggml IQ codebooks (tools/gen_iq_grids.py) */

#define QK_K 256

/* ggml type ids (public ggml spec; mirrored by gguf.constants.GGMLQuantizationType) */
enum {
    T_Q4_0 = 2, T_Q4_1 = 3, T_Q5_0 = 6, T_Q5_1 = 7, T_Q8_0 = 8,
    T_Q2_K = 10, T_Q3_K = 11, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14,
    T_IQ2_XXS = 16, T_IQ2_XS = 17, T_IQ3_XXS = 18, T_IQ1_S = 19, T_IQ3_S = 21, T_IQ2_S = 22, T_IQ1_M = 29, T_IQ4_NL = 20, T_IQ4_XS = 23, T_TQ1_0 = 34, T_TQ2_0 = 35
};

static const int8_t kvalues_iq4nl[16] = {-127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113};

/* exact IEEE half -> float */
static float h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal */
            int e = -1;
            do { e++; man <<= 1; } while ((man & 0x400u) == 0);
            man &= 0x3FFu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
static uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

int64_t orc_type_block_elems(int t) {
    switch (t) {
        case T_Q4_0: case T_Q4_1: case T_Q5_0: case T_Q5_1: case T_Q8_0: case T_IQ4_NL: return 32;
        case T_Q2_K: case T_Q3_K: case T_Q4_K: case T_Q5_K: case T_Q6_K: case T_IQ4_XS: case T_TQ1_0: case T_TQ2_0: case T_IQ2_XXS: case T_IQ2_XS: case T_IQ3_XXS: case T_IQ2_S: case T_IQ3_S: case T_IQ1_S: case T_IQ1_M: return QK_K;
        default: return 0;
    }
}
int64_t orc_type_block_bytes(int t) {
    switch (t) {
        case T_Q4_0: return 18; case T_Q4_1: return 20; case T_Q5_0: return 22; case T_Q5_1: return 24;
        case T_Q8_0: return 34; case T_Q2_K: return 84; case T_Q3_K: return 110; case T_Q4_K: return 144;
        case T_Q5_K: return 176; case T_Q6_K: return 210; case T_IQ4_NL: return 18; case T_IQ4_XS: return 136;
        case T_TQ1_0: return 54; case T_TQ2_0: return 66; case T_IQ2_XXS: return 66; case T_IQ2_XS: return 74; case T_IQ3_XXS: return 98;
        case T_IQ2_S: return 82; case T_IQ3_S: return 110; case T_IQ1_S: return 50; case T_IQ1_M: return 56;
        default: return 0;
    }
}
/* width of the scale sub-block (granularity of a_p, b_p and of the integer partials) */
int orc_type_sub(int t) {
    switch (t) {
        case T_Q2_K: case T_Q3_K: case T_Q6_K: return 16;
        case T_IQ2_XXS: case T_IQ2_XS: case T_IQ3_XXS: case T_IQ2_S: case T_IQ3_S: case T_IQ1_S: case T_IQ1_M: return 16;  /* reported at 16 (IQ2_XS's native granularity) for all grid formats */
        default: return 32;
    }
}

/* Q4_K / Q5_K 6-bit scale+min unpack */
static void k4_scale_min(int j, const uint8_t* s, int* sc, int* m) {
    if (j < 4) {
        *sc = s[j] & 63;
        *m = s[j + 4] & 63;
    } else {
        *sc = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4);
        *m = (s[j + 4] >> 4) | ((s[j] >> 6) << 4);
    }
}

/*
 * Decompose ONE block of type t into qi[block_elems], a[block_elems/sub], b[block_elems/sub].
 */
static void decompose_block(int t, const uint8_t* p, int8_t* qi, float* a, float* b) {
    switch (t) {
        case T_Q4_0: {
            float d = h2f(rd16(p));
            const uint8_t* qs = p + 2;
            for (int j = 0; j < 16; j++) { qi[j] = (int8_t)((qs[j] & 0xF) - 8); qi[j + 16] = (int8_t)((qs[j] >> 4) - 8); }
            a[0] = d; b[0] = 0.0f;
        } break;
        case T_Q4_1: {
            float d = h2f(rd16(p)), m = h2f(rd16(p + 2));
            const uint8_t* qs = p + 4;
            for (int j = 0; j < 16; j++) { qi[j] = (int8_t)(qs[j] & 0xF); qi[j + 16] = (int8_t)(qs[j] >> 4); }
            a[0] = d; b[0] = -m;
        } break;
        case T_Q5_0: {
            float d = h2f(rd16(p));
            uint32_t qh = rd32(p + 2);
            const uint8_t* qs = p + 6;
            for (int j = 0; j < 16; j++) {
                int h0 = (qh >> j) & 1, h1 = (qh >> (j + 16)) & 1;
                qi[j] = (int8_t)(((qs[j] & 0xF) | (h0 << 4)) - 16);
                qi[j + 16] = (int8_t)(((qs[j] >> 4) | (h1 << 4)) - 16);
            }
            a[0] = d; b[0] = 0.0f;
        } break;
        case T_Q5_1: {
            float d = h2f(rd16(p)), m = h2f(rd16(p + 2));
            uint32_t qh = rd32(p + 4);
            const uint8_t* qs = p + 8;
            for (int j = 0; j < 16; j++) {
                int h0 = (qh >> j) & 1, h1 = (qh >> (j + 16)) & 1;
                qi[j] = (int8_t)((qs[j] & 0xF) | (h0 << 4));
                qi[j + 16] = (int8_t)((qs[j] >> 4) | (h1 << 4));
            }
            a[0] = d; b[0] = -m;
        } break;
        case T_Q8_0: {
            float d = h2f(rd16(p));
            for (int j = 0; j < 32; j++) qi[j] = (int8_t)p[2 + j];
            a[0] = d; b[0] = 0.0f;
        } break;
        case T_IQ4_NL: {
            float d = h2f(rd16(p));
            const uint8_t* qs = p + 2;
            for (int j = 0; j < 16; j++) { qi[j] = kvalues_iq4nl[qs[j] & 0xF]; qi[j + 16] = kvalues_iq4nl[qs[j] >> 4]; }
            a[0] = d; b[0] = 0.0f;
        } break;
        case T_Q2_K: { /* [u8 scales[16]][u8 qs[64]][f16 d][f16 dmin] */
            const uint8_t* sc = p; const uint8_t* qs = p + 16;
            float d = h2f(rd16(p + 80)), dmin = h2f(rd16(p + 82));
            for (int s = 0; s < 16; s++) { a[s] = d * (float)(sc[s] & 0xF); b[s] = dmin * (float)(sc[s] >> 4); }
            /* element e: half n=e/128, shift group j=(e%128)/32, l=e%32 -> (qs[32n+l] >> 2j) & 3 */
            for (int e = 0; e < 256; e++) {
                int n = e / 128, j = (e % 128) / 32, l = e % 32;
                qi[e] = (int8_t)((qs[32 * n + l] >> (2 * j)) & 3);
            }
        } break;
        case T_Q3_K: { /* [u8 hmask[32]][u8 qs[64]][u8 scales[12]][f16 d] */
            const uint8_t* hm = p; const uint8_t* qs = p + 32; const uint8_t* s = p + 96;
            float d = h2f(rd16(p + 108));
            for (int i = 0; i < 16; i++) {
                int lo = (i < 8) ? (s[i] & 0xF) : (s[i - 8] >> 4);
                int hi = (s[8 + (i % 4)] >> (2 * (i / 4))) & 3;
                int scl = (lo | (hi << 4)) - 32;
                a[i] = d * (float)scl; b[i] = 0.0f;
            }
            for (int e = 0; e < 256; e++) {
                int n = e / 128, j = (e % 128) / 32, l = e % 32;
                int ql = (qs[32 * n + l] >> (2 * j)) & 3;
                int hbit = (hm[l] >> (4 * n + j)) & 1;
                qi[e] = (int8_t)(ql - (hbit ? 0 : 4));
            }
        } break;
        case T_Q4_K: { /* [f16 d][f16 dmin][u8 s[12]][u8 qs[128]] */
            float d = h2f(rd16(p)), dmin = h2f(rd16(p + 2));
            const uint8_t* s = p + 4; const uint8_t* qs = p + 16;
            for (int j = 0; j < 8; j++) { int sc, m; k4_scale_min(j, s, &sc, &m); a[j] = d * (float)sc; b[j] = dmin * (float)m; }
            for (int c = 0; c < 4; c++)
                for (int l = 0; l < 32; l++) { qi[64 * c + l] = (int8_t)(qs[32 * c + l] & 0xF); qi[64 * c + 32 + l] = (int8_t)(qs[32 * c + l] >> 4); }
        } break;
        case T_Q5_K: { /* [f16 d][f16 dmin][u8 s[12]][u8 qh[32]][u8 qs[128]] */
            float d = h2f(rd16(p)), dmin = h2f(rd16(p + 2));
            const uint8_t* s = p + 4; const uint8_t* qh = p + 16; const uint8_t* qs = p + 48;
            for (int j = 0; j < 8; j++) { int sc, m; k4_scale_min(j, s, &sc, &m); a[j] = d * (float)sc; b[j] = dmin * (float)m; }
            for (int c = 0; c < 4; c++)
                for (int l = 0; l < 32; l++) {
                    int h0 = (qh[l] >> (2 * c)) & 1, h1 = (qh[l] >> (2 * c + 1)) & 1;
                    qi[64 * c + l] = (int8_t)((qs[32 * c + l] & 0xF) | (h0 << 4));
                    qi[64 * c + 32 + l] = (int8_t)((qs[32 * c + l] >> 4) | (h1 << 4));
                }
        } break;
        case T_Q6_K: { /* [u8 ql[128]][u8 qh[64]][i8 sc[16]][f16 d] */
            const uint8_t* ql = p; const uint8_t* qh = p + 128; const int8_t* sc = (const int8_t*)(p + 192);
            float d = h2f(rd16(p + 208));
            for (int i = 0; i < 16; i++) { a[i] = d * (float)sc[i]; b[i] = 0.0f; }
            for (int h = 0; h < 2; h++)
                for (int l = 0; l < 32; l++) {
                    const uint8_t* L = ql + 64 * h; const uint8_t* H = qh + 32 * h;
                    int q1 = (L[l] & 15) | ((H[l] & 3) << 4);
                    int q2 = (L[l + 32] & 15) | (((H[l] >> 2) & 3) << 4);
                    int q3 = (L[l] >> 4) | (((H[l] >> 4) & 3) << 4);
                    int q4 = (L[l + 32] >> 4) | (((H[l] >> 6) & 3) << 4);
                    qi[128 * h + l] = (int8_t)(q1 - 32);
                    qi[128 * h + l + 32] = (int8_t)(q2 - 32);
                    qi[128 * h + l + 64] = (int8_t)(q3 - 32);
                    qi[128 * h + l + 96] = (int8_t)(q4 - 32);
                }
        } break;
        case T_IQ4_XS: { /* [f16 d][u16 scales_h][u8 scales_l[4]][u8 qs[128]] */
            float d = h2f(rd16(p));
            uint16_t sh = rd16(p + 2);
            const uint8_t* sl = p + 4; const uint8_t* qs = p + 8;
            for (int ib = 0; ib < 8; ib++) {
                int lo = (sl[ib / 2] >> (4 * (ib % 2))) & 0xF;
                int hi = (sh >> (2 * ib)) & 3;
                a[ib] = d * (float)((lo | (hi << 4)) - 32); b[ib] = 0.0f;
                for (int j = 0; j < 16; j++) {
                    qi[32 * ib + j] = kvalues_iq4nl[qs[16 * ib + j] & 0xF];
                    qi[32 * ib + 16 + j] = kvalues_iq4nl[qs[16 * ib + j] >> 4];
                }
            }
        } break;
        case T_IQ2_XXS: { /* [f16 d][8 x (u32 grid indices, u32 4 x 7-bit sign index | 4-bit scale << 28)]; gguf.quants.IQ2_XXS */
            float d = h2f(rd16(p));
            for (int ib = 0; ib < 8; ib++) {
                uint32_t q0 = rd32(p + 2 + 8 * ib), q1 = rd32(p + 2 + 8 * ib + 4);
                float db = (d * (0.5f + (float)(q1 >> 28))) * 0.25f;
                a[2 * ib] = a[2 * ib + 1] = db; b[2 * ib] = b[2 * ib + 1] = 0.0f;
                for (int l = 0; l < 4; l++) {
                    const uint8_t* g = iq2xxs_grid[(q0 >> (8 * l)) & 0xFF];
                    uint8_t sg = iq_ksigns[(q1 >> (7 * l)) & 127];
                    for (int j = 0; j < 8; j++) qi[32 * ib + 8 * l + j] = (int8_t)(((sg >> j) & 1) ? -(int)g[j] : (int)g[j]);
                }
            }
        } break;
        case T_IQ2_XS: { /* [f16 d][u16 qs[32]: 9-bit grid index | 7-bit sign index << 9][u8 scales[8]: two 4-bit scales, one per 16] */
            float d = h2f(rd16(p));
            const uint8_t* sc = p + 2 + 64;
            for (int sb = 0; sb < 16; sb++) {
                int s4 = (sc[sb / 2] >> (4 * (sb % 2))) & 0xF;
                a[sb] = (d * (0.5f + (float)s4)) * 0.25f; b[sb] = 0.0f;
            }
            for (int i = 0; i < 32; i++) {
                uint16_t q = rd16(p + 2 + 2 * i);
                const uint8_t* g = iq2xs_grid[q & 511];
                uint8_t sg = iq_ksigns[q >> 9];
                for (int j = 0; j < 8; j++) qi[8 * i + j] = (int8_t)(((sg >> j) & 1) ? -(int)g[j] : (int)g[j]);
            }
        } break;
        case T_IQ3_XXS: { /* [f16 d][u8 qs[64] grid indices (4 values each)][8 x u32: 4 x 7-bit sign index | 4-bit scale << 28] */
            float d = h2f(rd16(p));
            const uint8_t* qs = p + 2;
            for (int ib = 0; ib < 8; ib++) {
                uint32_t q1 = rd32(p + 2 + 64 + 4 * ib);
                float db = (d * (0.5f + (float)(q1 >> 28))) * 0.5f;
                a[2 * ib] = a[2 * ib + 1] = db; b[2 * ib] = b[2 * ib + 1] = 0.0f;
                for (int l = 0; l < 4; l++) {
                    uint8_t sg = iq_ksigns[(q1 >> (7 * l)) & 127];
                    for (int j = 0; j < 8; j++) {
                        int v = iq3xxs_grid[qs[8 * ib + 2 * l + (j >> 2)]][j & 3];
                        qi[32 * ib + 8 * l + j] = (int8_t)(((sg >> j) & 1) ? -v : v);
                    }
                }
            }
        } break;
        case T_IQ2_S: { /* [f16 d][u8 qs[32]][u8 signs[32]][u8 qh[8]][u8 scales[8]]: 10-bit grid index = qs | 2 bits of qh << 8 */
            float d = h2f(rd16(p));
            const uint8_t* qs = p + 2; const uint8_t* sg = p + 34; const uint8_t* qh = p + 66; const uint8_t* sc = p + 74;
            for (int sb = 0; sb < 16; sb++) {
                int s4 = (sc[sb / 2] >> (4 * (sb % 2))) & 0xF;
                a[sb] = (d * (0.5f + (float)s4)) * 0.25f; b[sb] = 0.0f;
            }
            for (int i = 0; i < 32; i++) {
                int idx = qs[i] | (((qh[i / 4] >> (2 * (i % 4))) & 3) << 8);
                const uint8_t* g = iq2s_grid[idx];
                for (int j = 0; j < 8; j++) qi[8 * i + j] = (int8_t)(((sg[i] >> j) & 1) ? -(int)g[j] : (int)g[j]);
            }
        } break;
        case T_IQ3_S: { /* [f16 d][u8 qs[64]][u8 qh[8]][u8 signs[32]][u8 scales[4]]: 9-bit grid index = qs | 1 bit of qh << 8 */
            float d = h2f(rd16(p));
            const uint8_t* qs = p + 2; const uint8_t* qh = p + 66; const uint8_t* sg = p + 74; const uint8_t* sc = p + 106;
            for (int ib = 0; ib < 8; ib++) {
                int s4 = (sc[ib / 2] >> (4 * (ib % 2))) & 0xF;
                a[2 * ib] = a[2 * ib + 1] = d * (float)(1 + 2 * s4); b[2 * ib] = b[2 * ib + 1] = 0.0f;
            }
            for (int i = 0; i < 64; i++) {
                int idx = qs[i] | (((qh[i / 8] >> (i % 8)) & 1) << 8);
                const uint8_t* g = iq3s_grid[idx];
                for (int j = 0; j < 4; j++) {
                    int e = 4 * i + j;
                    qi[e] = (int8_t)(((sg[e / 8] >> (e % 8)) & 1) ? -(int)g[j] : (int)g[j]);
                }
            }
        } break;
        case T_IQ1_S: { /* [f16 d][u8 qs[32]][u16 qh[8]]: per 32: 3-bit scale (bits 12-14), delta sign (bit 15), 4 x 11-bit grid index;
                           w = d (2 s + 1) (g + delta), delta = +-1/8  ==>  integer form: v = 8 g +- 1, a = d (2 s + 1) / 8 (all exact) */
            float d = h2f(rd16(p));
            const uint8_t* qs = p + 2;
            for (int ib = 0; ib < 8; ib++) {
                uint16_t qh = rd16(p + 34 + 2 * ib);
                float dl = d * (float)(2 * ((qh >> 12) & 7) + 1);
                a[2 * ib] = a[2 * ib + 1] = dl * 0.125f; b[2 * ib] = b[2 * ib + 1] = 0.0f;
                int dsign = (qh & 0x8000) ? -1 : 1;
                for (int l = 0; l < 4; l++) {
                    const int8_t* g = iq1s_grid[qs[4 * ib + l] | (((qh >> (3 * l)) & 7) << 8)];
                    for (int j = 0; j < 8; j++) qi[32 * ib + 8 * l + j] = (int8_t)(8 * g[j] + dsign);
                }
            }
        } break;
        case T_IQ1_M: { /* [u8 qs[32]][u8 qh[16]][u16 scales[4]]: f16 d scattered over the top nibbles of the scale words, a 3-bit scale
                           per 16, per 8 elements an 11-bit grid index (qs | 3 bits of a qh nibble << 8) and a delta sign (bit 3 of the nibble) */
            const uint8_t* qs = p; const uint8_t* qh = p + 32;
            uint16_t sc[4];
            for (int i = 0; i < 4; i++) sc[i] = rd16(p + 48 + 2 * i);
            uint16_t dbits = (uint16_t)((sc[0] >> 12) | ((sc[1] >> 8) & 0x00F0) | ((sc[2] >> 4) & 0x0F00) | (sc[3] & 0xF000));
            float d = h2f(dbits);
            for (int sb = 0; sb < 16; sb++) {
                int s3 = (sc[sb / 4] >> (3 * (sb % 4))) & 7;
                a[sb] = (d * (float)(2 * s3 + 1)) * 0.125f; b[sb] = 0.0f;
            }
            for (int i = 0; i < 32; i++) {
                int nib = (qh[i / 2] >> (4 * (i % 2))) & 0xF;
                const int8_t* g = iq1s_grid[qs[i] | ((nib & 7) << 8)];
                int dsign = (nib & 8) ? -1 : 1;
                for (int j = 0; j < 8; j++) qi[8 * i + j] = (int8_t)(8 * g[j] + dsign);
            }
        } break;
        case T_TQ2_0: { /* ternary, 2 bits: [u8 qs[64]][f16 d]; element 128 n + 32 l + m = ((qs[32 n + m] >> 2 l) & 3) - 1 */
            float d = h2f(rd16(p + 64));
            for (int sb = 0; sb < 8; sb++) { a[sb] = d; b[sb] = 0.0f; }
            for (int e = 0; e < 256; e++) {
                int n = e / 128, l = (e % 128) / 32, m = e % 32;
                qi[e] = (int8_t)(((p[32 * n + m] >> (2 * l)) & 3) - 1);
            }
        } break;
        case T_TQ1_0: { /* ternary, base 3: [u8 qs[48]][u8 qh[4]][f16 d]; 5 trits per qs byte, 4 per qh byte:
                           trit n of byte x = ((uint8)(x * 3^n) * 3) >> 8 */
            static const uint8_t pow3[5] = {1, 3, 9, 27, 81};
            float d = h2f(rd16(p + 52));
            for (int sb = 0; sb < 8; sb++) { a[sb] = d; b[sb] = 0.0f; }
            for (int n = 0; n < 5; n++)
                for (int m = 0; m < 32; m++) qi[32 * n + m] = (int8_t)((((uint16_t)(uint8_t)(p[m] * pow3[n])) * 3 >> 8) - 1);
            for (int n = 0; n < 5; n++)
                for (int m = 0; m < 16; m++) qi[160 + 16 * n + m] = (int8_t)((((uint16_t)(uint8_t)(p[32 + m] * pow3[n])) * 3 >> 8) - 1);
            for (int n = 0; n < 4; n++)
                for (int m = 0; m < 4; m++) qi[240 + 4 * n + m] = (int8_t)((((uint16_t)(uint8_t)(p[48 + m] * pow3[n])) * 3 >> 8) - 1);
        } break;
        default: break;
    }
}

/* W = a*qi - b per element, separate multiply and subtract (bit-exact contract #1) */
void orc_decompose_ggml(int t, const uint8_t* blocks, int64_t nelem, int8_t* qi, float* a, float* b) {
    int64_t be = orc_type_block_elems(t), bb = orc_type_block_bytes(t);
    int sub = orc_type_sub(t);
    int64_t nb = nelem / be;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nb; i++)
        decompose_block(t, blocks + i * bb, qi + i * be, a + i * (be / sub), b + i * (be / sub));
}

void orc_dequant_ggml(int t, const uint8_t* blocks, int64_t nelem, float* out) {
    int64_t be = orc_type_block_elems(t), bb = orc_type_block_bytes(t);
    int sub = orc_type_sub(t);
    int64_t nb = nelem / be;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nb; i++) {
        int8_t qi[QK_K]; float a[16], b[16];
        decompose_block(t, blocks + i * bb, qi, a, b);
        float* o = out + i * be;
        for (int e = 0; e < be; e++) {
            float prod = a[e / sub] * (float)qi[e];
            o[e] = prod - b[e / sub];
        }
    }
}

/*
 * Activation quantizer, per 32-element block ("Q8_1 style", the convention of dp4a matvec
 * kernels; boostr's CUDA side is described as dp4a kernels at reference README.md:120):
 *   d = amax / 127;  id = d ? 1/d : 0;  q = roundf(x * id)   (IEEE f32 ops, round half away) -- the ggml
 *   quantize_row_q8_0 recipe, bit-exact with gguf.quants' Q8_0 quantizer (tests/golden)
 * bsum16[j] = sum of q over 16-element half blocks (int32).
 */
void orc_quantize_act(const float* x, int64_t M, int64_t K, int8_t* q, float* d, int32_t* bsum16) {
    int64_t nb = K / 32;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M * nb; i++) {
        const float* xb = x + i * 32;
        float amax = 0.0f;
        for (int j = 0; j < 32; j++) { float v = fabsf(xb[j]); if (v > amax) amax = v; }
        float dd = amax / 127.0f;
        float id = (dd != 0.0f) ? 1.0f / dd : 0.0f;
        d[i] = dd;
        int32_t s0 = 0, s1 = 0;
        for (int j = 0; j < 32; j++) {
            int8_t v = (int8_t)roundf(xb[j] * id);
            q[i * 32 + j] = v;
            if (j < 16) s0 += v; else s1 += v;
        }
        bsum16[2 * i] = s0; bsum16[2 * i + 1] = s1;
    }
}

/* integer partials: out[m][n][p] = sum_{k in sub-block p} qi[n][k] * xq[m][k]   (contract #2) */
void orc_int_partials(const int8_t* qi, const int8_t* xq, int sub, int64_t N, int64_t K, int64_t M, int32_t* out) {
    int64_t P = K / sub;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; n++)
        for (int64_t m = 0; m < M; m++)
            for (int64_t p = 0; p < P; p++) {
                int32_t s = 0;
                for (int e = 0; e < sub; e++) s += (int32_t)qi[n * K + p * sub + e] * (int32_t)xq[m * K + p * sub + e];
                out[(m * N + n) * P + p] = s;
            }
}

/* flavour A from a dense dequantized matrix: Y[m,n] = sum_k X[m,k] W[n,k] (+bias), double accum */
void orc_matmul_dense(const float* W, const float* X, const float* bias, int64_t N, int64_t K, int64_t M, float* Y) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; n++)
        for (int64_t m = 0; m < M; m++) {
            double acc = 0.0;
            const float* w = W + n * K; const float* x = X + m * K;
            for (int64_t k = 0; k < K; k++) acc += (double)w[k] * (double)x[k];
            if (bias) acc += (double)bias[n];
            Y[m * N + n] = (float)acc;
        }
}

/* flavour A straight from ggml blocks (row-wise dequant, no dense copy) */
void orc_matmul_ggml_f32(int t, const uint8_t* blocks, int64_t N, int64_t K, const float* X, int64_t M, float* Y) {
    int64_t be = orc_type_block_elems(t), bb = orc_type_block_bytes(t);
    int64_t row_bytes = K / be * bb;
#pragma omp parallel
    {
        float* w = (float*)malloc(sizeof(float) * (size_t)K);
#pragma omp for schedule(static)
        for (int64_t n = 0; n < N; n++) {
            int sub = orc_type_sub(t);
            for (int64_t i = 0; i < K / be; i++) {
                int8_t qi[QK_K]; float a[16], b[16];
                decompose_block(t, blocks + n * row_bytes + i * bb, qi, a, b);
                for (int e = 0; e < be; e++) { float prod = a[e / sub] * (float)qi[e]; w[i * be + e] = prod - b[e / sub]; }
            }
            for (int64_t m = 0; m < M; m++) {
                double acc = 0.0;
                const float* x = X + m * K;
                for (int64_t k = 0; k < K; k++) acc += (double)w[k] * (double)x[k];
                Y[m * N + n] = (float)acc;
            }
        }
        free(w);
    }
}

/*
 * flavour B from the decomposition: int8 activations (orc_quantize_act), integer dot per sub-block,
 *   Y[m,n] = sum_p (a_p * dx_blk(p)) * partial_p  -  sum_p (b_p * dx_blk(p)) * bsum_p   (+bias)
 * accumulated in double (exact products, see below): bit-exact contract #3 for the dp4a decode path.
 */
void orc_matmul_q8(const int8_t* qi, const float* a, const float* b, int sub, int64_t N, int64_t K,
                   const int8_t* xq, const float* xd, const int32_t* xbsum16, int64_t M, const float* bias, float* Y) {
    int64_t P = K / sub;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; n++)
        for (int64_t m = 0; m < M; m++) {
            double acc = 0.0;
            for (int64_t p = 0; p < P; p++) {
                int32_t s = 0;
                const int8_t* wq = qi + n * K + p * sub; const int8_t* xx = xq + m * K + p * sub;
                for (int e = 0; e < sub; e++) s += (int32_t)wq[e] * (int32_t)xx[e];
                int64_t blk = (p * sub) / 32;
                float dx = xd[m * (K / 32) + blk];
                int32_t bs;
                if (sub == 32) bs = xbsum16[(m * (K / 32) + blk) * 2] + xbsum16[(m * (K / 32) + blk) * 2 + 1];
                else bs = xbsum16[m * (K / 16) + p];
                /* each product is exact in f64 (24-bit x <=24-bit), so the f64 sum is order independent to 1e-16:
                 * the GPU kernel forms the very same terms and is bit-comparable after the final f32 rounding */
                acc += (double)(a[n * P + p] * dx) * (double)s;
                acc -= (double)(b[n * P + p] * dx) * (double)bs;
            }
            if (bias) acc += (double)bias[n];
            Y[m * N + n] = (float)acc;
        }
}

/*
 * Packed-block flavour B matvec used as the *CPU baseline* (bench.py cpu_baseline / --impl reference):
 * streams the ggml blocks once per call like a CPU inference path does, OpenMP over rows.
 */
void orc_matvec_ggml_q8(int t, const uint8_t* blocks, int64_t N, int64_t K,
                        const int8_t* xq, const float* xd, const int32_t* xbsum16, int64_t M, float* Y) {
    int64_t be = orc_type_block_elems(t), bb = orc_type_block_bytes(t);
    int sub = orc_type_sub(t);
    int64_t row_bytes = K / be * bb;
    int spb = (int)(be / sub);
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; n++) {
        float acc[64];
        for (int64_t m = 0; m < M && m < 64; m++) acc[m] = 0.0f;
        for (int64_t i = 0; i < K / be; i++) {
            int8_t qi[QK_K]; float a[16], b[16];
            decompose_block(t, blocks + n * row_bytes + i * bb, qi, a, b);
            for (int64_t m = 0; m < M && m < 64; m++) {
                const int8_t* xx = xq + m * K + i * be;
                for (int p = 0; p < spb; p++) {
                    int32_t s = 0;
                    for (int e = 0; e < sub; e++) s += (int32_t)qi[p * sub + e] * (int32_t)xx[p * sub + e];
                    int64_t blk = (i * be + p * sub) / 32;
                    float dx = xd[m * (K / 32) + blk];
                    int32_t bs;
                    if (sub == 32) bs = xbsum16[(m * (K / 32) + blk) * 2] + xbsum16[(m * (K / 32) + blk) * 2 + 1];
                    else bs = xbsum16[m * (K / 16) + (i * be) / 16 + p];
                    acc[m] += (a[p] * dx) * (float)s - (b[p] * dx) * (float)bs;
                }
            }
        }
        for (int64_t m = 0; m < M && m < 64; m++) Y[m * N + n] = acc[m];
    }
}

/* ------------------------------------------------------------------------------------------- */
/* AWQ (reference src/loader/safetensors/awq.rs)                                               */
/* ------------------------------------------------------------------------------------------- */
static const uint32_t AWQ_SHIFTS[8] = {0, 16, 4, 20, 8, 24, 12, 28}; /* awq.rs:29-32 */

/* awq.rs:242-263: packed qzeros [G, N/8] u32 -> f32 [G, N] */
void orc_awq_unpack_zeros(const uint32_t* packed, int64_t G, int64_t N, float* out) {
    int64_t n8 = N / 8;
    for (int64_t g = 0; g < G; g++)
        for (int64_t j = 0; j < n8; j++) {
            uint32_t v = packed[g * n8 + j];
            for (int k = 0; k < 8; k++) out[g * N + j * 8 + k] = (float)((v >> AWQ_SHIFTS[k]) & 0xF);
        }
}

/* awq.rs:190-226: qweight u32 [K, N/8], scales f32 [K/gs, N], zeros f32 [K/gs, N]; logical [N, K].
 * W[n,k] = (q(k,n) - z[k/gs,n]) * s[k/gs,n],  q(k,n) = (qweight[k,n/8] >> AWQ_SHIFTS[n%8]) & 0xF */
void orc_awq_dequant(const uint32_t* qweight, const float* scales, const float* zeros, int64_t gs,
                     int64_t N, int64_t K, float* out) {
    int64_t n8 = N / 8;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; n++)
        for (int64_t k = 0; k < K; k++) {
            uint32_t q = (qweight[k * n8 + n / 8] >> AWQ_SHIFTS[n % 8]) & 0xF;
            float diff = (float)q - zeros[(k / gs) * N + n];
            out[n * K + k] = diff * scales[(k / gs) * N + n];
        }
}

/* decomposition for AWQ: qi = q - z (integer zero point), a = s per 32-block, b = 0 */
void orc_awq_decompose(const uint32_t* qweight, const float* scales, const float* zeros, int64_t gs,
                       int64_t N, int64_t K, int8_t* qi, float* a, float* b) {
    int64_t n8 = N / 8;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; n++)
        for (int64_t k = 0; k < K; k++) {
            int q = (int)((qweight[k * n8 + n / 8] >> AWQ_SHIFTS[n % 8]) & 0xF);
            int z = (int)zeros[(k / gs) * N + n];
            qi[n * K + k] = (int8_t)(q - z);
            if (k % 32 == 0) { a[n * (K / 32) + k / 32] = scales[(k / gs) * N + n]; b[n * (K / 32) + k / 32] = 0.0f; }
        }
}

/* ------------------------------------------------------------------------------------------- */
/* GPTQ (reference src/loader/safetensors/gptq.rs)                                             */
/* ------------------------------------------------------------------------------------------- */
/* gptq.rs:198-259: qweight u32 [K/8, N] (sequential 4-bit along K), scales f32 [G, N],
 * qzeros u32 [G, N/8] kept packed (sequential 4-bit along N), g_idx i32 [K] (nullable).
 * W[n,k] = (q(k,n) - zp(g,n)) * s[g,n];  g = g_idx ? g_idx[k] : k/gs;  zp = z + zero_plus_one
 * (zero_plus_one is NOT observable from blazr: AutoGPTQ v1 files store z-1; caller chooses). */
static inline int gptq_q(const uint32_t* qweight, int64_t N, int64_t k, int64_t n) {
    return (int)((qweight[(k / 8) * N + n] >> (4 * (k % 8))) & 0xF);
}
static inline int gptq_z(const uint32_t* qzeros, int64_t N, int64_t g, int64_t n) {
    return (int)((qzeros[g * (N / 8) + n / 8] >> (4 * (n % 8))) & 0xF);
}
void orc_gptq_dequant(const uint32_t* qweight, const float* scales, const uint32_t* qzeros, const int32_t* g_idx,
                      int64_t gs, int zero_plus_one, int64_t N, int64_t K, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; n++)
        for (int64_t k = 0; k < K; k++) {
            int64_t g = g_idx ? g_idx[k] : k / gs;
            int q = gptq_q(qweight, N, k, n);
            int zp = gptq_z(qzeros, N, g, n) + zero_plus_one;
            float diff = (float)(q - zp);
            out[n * K + k] = diff * scales[g * N + n];
        }
}

/* Stable permutation that makes groups contiguous: perm[k'] = k, sorted by (g_idx[k], k).
 * Returns 0 when every group has exactly gs members (required), -1 otherwise. */
int orc_gptq_perm(const int32_t* g_idx, int64_t gs, int64_t K, int32_t* perm) {
    int64_t G = K / gs;
    int64_t* cnt = (int64_t*)calloc((size_t)G + 1, sizeof(int64_t));
    for (int64_t k = 0; k < K; k++) { if (g_idx[k] < 0 || g_idx[k] >= G) { free(cnt); return -1; } cnt[g_idx[k] + 1]++; }
    for (int64_t g = 0; g < G; g++) { if (cnt[g + 1] != gs) { free(cnt); return -1; } }
    for (int64_t g = 0; g < G; g++) cnt[g + 1] += cnt[g];
    for (int64_t k = 0; k < K; k++) perm[cnt[g_idx[k]]++] = (int32_t)k;
    free(cnt);
    return 0;
}

/* decomposition in PERMUTED k' order (identity when g_idx == NULL): qi[n,k'] = q(perm[k'],n) - zp, a per 32 */
void orc_gptq_decompose(const uint32_t* qweight, const float* scales, const uint32_t* qzeros, const int32_t* perm,
                        int64_t gs, int zero_plus_one, int64_t N, int64_t K, int8_t* qi, float* a, float* b) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; n++)
        for (int64_t kp = 0; kp < K; kp++) {
            int64_t k = perm ? perm[kp] : kp;
            int64_t g = kp / gs;
            int q = gptq_q(qweight, N, k, n);
            int zp = gptq_z(qzeros, N, g, n) + zero_plus_one;
            qi[n * K + kp] = (int8_t)(q - zp);
            if (kp % 32 == 0) { a[n * (K / 32) + kp / 32] = scales[g * N + n]; b[n * (K / 32) + kp / 32] = 0.0f; }
        }
}

/* ------------------------------------------------------------------------------------------- */
/* TP shard rule (reference src/engine/tensor_parallel.rs:61-67): even split, remainder to low ranks */
/* ------------------------------------------------------------------------------------------- */
void orc_shard_range(int64_t total, int64_t rank, int64_t world, int64_t* start, int64_t* end) {
    int64_t per = total / world, rem = total % world;
    *start = rank * per + (rank < rem ? rank : rem);
    *end = *start + per + (rank < rem ? 1 : 0);
}

/* f32 <-> bf16/f16 helpers so tests can build inputs identically to the device path */
float orc_h2f(uint16_t h) { return h2f(h); }

/* Deterministic expf, line-for-line the same IEEE operations as blazr_b200/csrc/common.cuh det_expf
 * (fmaf is correctly rounded on both sides; compile with -mfma so it is one instruction). */
static float det_expf(float x) {
    x = fminf(fmaxf(x, -87.0f), 88.0f);
    const float n = rintf(x * 1.44269504f);
    float r = fmaf(n, -0.693145752f, x);
    r = fmaf(n, -1.42860677e-6f, r);
    float p = 1.0f / 720.0f;
    p = fmaf(p, r, 1.0f / 120.0f);
    p = fmaf(p, r, 1.0f / 24.0f);
    p = fmaf(p, r, 1.0f / 6.0f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    union { uint32_t u; float f; } sc;
    sc.u = (uint32_t)(((int)n + 127) << 23);
    return p * sc.f;
}
void orc_det_expf(const float* x, int64_t n, float* out) {
    for (int64_t i = 0; i < n; i++) out[i] = det_expf(x[i]);
}
