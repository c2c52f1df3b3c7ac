/*
 * cpu_fast.c -- the TIMED CPU baseline (bench.py cpu_baseline / --impl reference): AVX2 + OpenMP kernels for the three
 * headline formats, written the way a tuned CPU inference path (ggml-style) computes a quantized matvec: weights stay
 * packed, activations are int8 with one f32 scale per 32 (orc_quantize_act), integer dot products per sub-block with
 * maddubs/madd, f32 accumulation.  TEST INFRASTRUCTURE / BASELINE ONLY -- the product never links this.
 * PARITY UNPINNED like the rest of the oracle (the reference's CPU kernels live in the un-vendored boostr/numr crates);
 * tests/test_oracle_golden.py holds it to the exact-accumulation oracle (orc_matmul_q8) within 1e-5.
 * Block layouts: public ggml spec.
 */
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

static inline float h2f_fast(uint16_t h) { return _cvtsh_ss(h); }

static inline float hsum8(__m256 v) {
    __m128 lo = _mm256_castps256_ps128(v), hi = _mm256_extractf128_ps(v, 1);
    lo = _mm_add_ps(lo, hi);
    lo = _mm_add_ps(lo, _mm_movehl_ps(lo, lo));
    lo = _mm_add_ss(lo, _mm_shuffle_ps(lo, lo, 1));
    return _mm_cvtss_f32(lo);
}
/* 32 unsigned bytes x 32 signed bytes -> 8 x i32 partial sums */
static inline __m256i dot_u8s8(__m256i u, __m256i s) {
    return _mm256_madd_epi16(_mm256_maddubs_epi16(u, s), _mm256_set1_epi16(1));
}
/* 32 signed x 32 signed */
static inline __m256i dot_s8s8(__m256i a, __m256i b) {
    return dot_u8s8(_mm256_sign_epi8(a, a), _mm256_sign_epi8(b, a));
}

static void q4k_scales(const uint8_t* s, int* sc, int* m) {
    for (int j = 0; j < 8; j++) {
        if (j < 4) { sc[j] = s[j] & 63; m[j] = s[j + 4] & 63; }
        else { sc[j] = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4); m[j] = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
    }
}

/* Y[m, n] = sum_k W[n, k] x[m, k]; xq int8 [M, K], xd f32 [M, K/32], xbsum16 i32 [M, K/16].  Returns 0 when the type
 * has a fast kernel, -1 otherwise (caller falls back to the generic port). */
int orc_matvec_fast(int t, const uint8_t* blocks, int64_t N, int64_t K, const int8_t* xq, const float* xd, const int32_t* xbsum16,
                    int64_t M, float* Y) {
    if (t != 8 && t != 12 && t != 14) return -1;
    if (t == 8) { /* Q8_0: [f16 d][i8 q[32]] */
        const int64_t nb = K / 32, row_bytes = nb * 34;
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < N; n++) {
            const uint8_t* row = blocks + n * row_bytes;
            for (int64_t m = 0; m < M; m++) {
                __m256 acc = _mm256_setzero_ps();
                const int8_t* xx = xq + m * K;
                const float* dx = xd + m * nb;
                for (int64_t i = 0; i < nb; i++) {
                    const uint8_t* b = row + i * 34;
                    uint16_t dh; memcpy(&dh, b, 2);
                    const __m256i q = _mm256_loadu_si256((const __m256i*)(b + 2));
                    const __m256i x = _mm256_loadu_si256((const __m256i*)(xx + 32 * i));
                    acc = _mm256_fmadd_ps(_mm256_cvtepi32_ps(dot_s8s8(q, x)), _mm256_set1_ps(h2f_fast(dh) * dx[i]), acc);
                }
                Y[m * N + n] = hsum8(acc);
            }
        }
        return 0;
    }
    if (t == 12) { /* Q4_K: [f16 d][f16 dmin][u8 s[12]][u8 qs[128]] */
        const int64_t nb = K / 256, row_bytes = nb * 144;
        const __m256i m4 = _mm256_set1_epi8(0x0F);
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < N; n++) {
            const uint8_t* row = blocks + n * row_bytes;
            for (int64_t m = 0; m < M; m++) {
                __m256 acc = _mm256_setzero_ps();
                float accm = 0.0f;
                const int8_t* xx = xq + m * K;
                const float* dx = xd + m * (K / 32);
                const int32_t* bs = xbsum16 + m * (K / 16);
                for (int64_t i = 0; i < nb; i++) {
                    const uint8_t* b = row + i * 144;
                    uint16_t dh, mh; memcpy(&dh, b, 2); memcpy(&mh, b + 2, 2);
                    const float d = h2f_fast(dh), dmin = h2f_fast(mh);
                    int sc[8], mn[8];
                    q4k_scales(b + 4, sc, mn);
                    const uint8_t* qs = b + 16;
                    for (int c = 0; c < 4; c++) {
                        const __m256i w = _mm256_loadu_si256((const __m256i*)(qs + 32 * c));
                        const __m256i lo = _mm256_and_si256(w, m4), hi = _mm256_and_si256(_mm256_srli_epi16(w, 4), m4);
                        const int64_t blk = i * 8 + 2 * c;
                        const __m256i x0 = _mm256_loadu_si256((const __m256i*)(xx + 32 * blk));
                        const __m256i x1 = _mm256_loadu_si256((const __m256i*)(xx + 32 * (blk + 1)));
                        acc = _mm256_fmadd_ps(_mm256_cvtepi32_ps(dot_u8s8(lo, x0)), _mm256_set1_ps(d * (float)sc[2 * c] * dx[blk]), acc);
                        acc = _mm256_fmadd_ps(_mm256_cvtepi32_ps(dot_u8s8(hi, x1)), _mm256_set1_ps(d * (float)sc[2 * c + 1] * dx[blk + 1]), acc);
                        accm += dmin * (float)mn[2 * c] * dx[blk] * (float)(bs[2 * blk] + bs[2 * blk + 1]);
                        accm += dmin * (float)mn[2 * c + 1] * dx[blk + 1] * (float)(bs[2 * blk + 2] + bs[2 * blk + 3]);
                    }
                }
                Y[m * N + n] = hsum8(acc) - accm;
            }
        }
        return 0;
    }
    /* Q6_K: [u8 ql[128]][u8 qh[64]][i8 sc[16]][f16 d]; 16 sub-blocks of 16; value = q6 - 32 */
    {
        const int64_t nb = K / 256, row_bytes = nb * 210;
        const __m256i m4 = _mm256_set1_epi8(0x0F), m2 = _mm256_set1_epi8(0x03);
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < N; n++) {
            const uint8_t* row = blocks + n * row_bytes;
            for (int64_t m = 0; m < M; m++) {
                __m256 acc = _mm256_setzero_ps();
                float accm = 0.0f;
                const int8_t* xx = xq + m * K;
                const float* dx = xd + m * (K / 32);
                const int32_t* bs = xbsum16 + m * (K / 16);
                for (int64_t i = 0; i < nb; i++) {
                    const uint8_t* b = row + i * 210;
                    const int8_t* sc = (const int8_t*)(b + 192);
                    uint16_t dh; memcpy(&dh, b + 208, 2);
                    const float d = h2f_fast(dh);
                    for (int h = 0; h < 2; h++) {
                        const __m256i l0 = _mm256_loadu_si256((const __m256i*)(b + 64 * h));
                        const __m256i l1 = _mm256_loadu_si256((const __m256i*)(b + 64 * h + 32));
                        const __m256i hh = _mm256_loadu_si256((const __m256i*)(b + 128 + 32 * h));
                        __m256i q[4];
                        q[0] = _mm256_or_si256(_mm256_and_si256(l0, m4), _mm256_slli_epi16(_mm256_and_si256(hh, m2), 4));
                        q[1] = _mm256_or_si256(_mm256_and_si256(l1, m4), _mm256_slli_epi16(_mm256_and_si256(_mm256_srli_epi16(hh, 2), m2), 4));
                        q[2] = _mm256_or_si256(_mm256_and_si256(_mm256_srli_epi16(l0, 4), m4), _mm256_slli_epi16(_mm256_and_si256(_mm256_srli_epi16(hh, 4), m2), 4));
                        q[3] = _mm256_or_si256(_mm256_and_si256(_mm256_srli_epi16(l1, 4), m4), _mm256_slli_epi16(_mm256_and_si256(_mm256_srli_epi16(hh, 6), m2), 4));
                        for (int r = 0; r < 4; r++) {
                            /* elements 128 h + 32 r + [0,32): two sub-blocks of 16 (lanes 0-3 / 4-7 after madd), one activation block */
                            const int64_t e0 = i * 256 + 128 * h + 32 * r;
                            const int64_t p = e0 / 16;
                            const __m256i x = _mm256_loadu_si256((const __m256i*)(xx + e0));
                            const __m256i s = dot_u8s8(q[r], x);
                            const float dxa = d * dx[e0 / 32];
                            const float a0 = dxa * (float)sc[8 * h + 2 * r], a1 = dxa * (float)sc[8 * h + 2 * r + 1];
                            acc = _mm256_fmadd_ps(_mm256_cvtepi32_ps(s), _mm256_set_m128(_mm_set1_ps(a1), _mm_set1_ps(a0)), acc);
                            accm += 32.0f * (a0 * (float)bs[p] + a1 * (float)bs[p + 1]);
                        }
                    }
                }
                Y[m * N + n] = hsum8(acc) - accm;
            }
        }
        return 0;
    }
}
