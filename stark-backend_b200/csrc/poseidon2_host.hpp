// Host-side Poseidon2 permutation for the Fiat–Shamir transcript (transcript.hpp): the prover hashes on the
// host between the ~450 sumcheck round kernels of a proof (2 permutations per round, 256+ for the column
// openings), and the scalar permutation shared with the kernels takes 2.6 us there — about 3 ms per proof.
// This AVX2 version keeps the 16-word state in two 256-bit registers (one M4 block per 128-bit lane, so the
// external layer is in-lane shuffles), does the 16 S-boxes of an external round as two 8-lane Montgomery
// chains and the internal diagonal as one vector Montgomery product.  Same function as p2::permute
// (bit-exact, canonical Montgomery words in and out); selected at run time when the CPU has AVX2.
//
// Replaces (reference, relative to /root/reference): the host DuplexChallenger's permutation,
// crates/stark-backend/src/transcript/duplex_sponge.rs:60-83 over p3-baby-bear's Poseidon2 (which has its own
// AVX2 backend in p3-monty-31).
#pragma once
#include <cstdint>

#include "poseidon2.cuh"

#if !defined(__CUDA_ARCH__) && (defined(__x86_64__) || defined(_M_X64))
#include <immintrin.h>
#define SWIRL_P2_HOST_AVX2 1
#endif

namespace p2host {

#ifdef SWIRL_P2_HOST_AVX2
#define P2H_TARGET __attribute__((target("avx2")))

P2H_TARGET static inline __m256i csub(__m256i x) {  // [0, 2p) -> [0, p)
    return _mm256_min_epu32(x, _mm256_sub_epi32(x, _mm256_set1_epi32((int)bb::P)));
}
P2H_TARGET static inline __m256i add(__m256i a, __m256i b) { return csub(_mm256_add_epi32(a, b)); }
// eight canonical Montgomery products
P2H_TARGET static inline __m256i mmul(__m256i a, __m256i b) {
    const __m256i P64 = _mm256_set1_epi64x((long long)bb::P), M64 = _mm256_set1_epi64x((long long)bb::NEG_PINV);
    const __m256i ao = _mm256_srli_epi64(a, 32), bo = _mm256_srli_epi64(b, 32);
    const __m256i pe = _mm256_mul_epu32(a, b), po = _mm256_mul_epu32(ao, bo);
    const __m256i qe = _mm256_mul_epu32(pe, M64), qo = _mm256_mul_epu32(po, M64);
    const __m256i se = _mm256_add_epi64(pe, _mm256_mul_epu32(qe, P64)), so = _mm256_add_epi64(po, _mm256_mul_epu32(qo, P64));
    return csub(_mm256_blend_epi32(_mm256_srli_epi64(se, 32), so, 0xAA));
}
P2H_TARGET static inline __m256i sbox7(__m256i x) {
    const __m256i x2 = mmul(x, x), x3 = mmul(x2, x), x4 = mmul(x2, x2);
    return mmul(x3, x4);
}
// y_i = 2 x_i + 3 x_{i+1} + x_{i+2} + x_{i+3} inside every 4-block (M4 is the circulant (2, 3, 1, 1))
P2H_TARGET static inline __m256i m4(__m256i x) {
    const __m256i r1 = _mm256_shuffle_epi32(x, 0x39);  // x_{i+1}
    const __m256i t = add(x, r1);
    const __m256i all = add(t, _mm256_shuffle_epi32(t, 0x4E));
    return add(add(all, x), add(r1, r1));
}
P2H_TARGET static inline void external_linear(__m256i& v0, __m256i& v1) {
    v0 = m4(v0);
    v1 = m4(v1);
    const __m256i s = add(v0, v1);
    const __m256i t = add(s, _mm256_permute2x128_si256(s, s, 0x01));  // column sums in both halves
    v0 = add(v0, t);
    v1 = add(v1, t);
}

static inline uint32_t scalar_sbox7(uint32_t x) {
    const uint32_t x2 = bb::mul(x, x), x3 = bb::mul(x2, x), x4 = bb::mul(x2, x2);
    return bb::mul(x3, x4);
}

P2H_TARGET static inline void permute_avx2(uint32_t s[16]) {
    // internal diagonal (-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 2^-8, 1/4, 1/8, 2^-27, -2^-8, -1/16, -2^-27), Montgomery form
    static const uint32_t DIAG[16] = {
        bb::mont_neg(2), bb::mont(1), bb::mont(2), bb::halve(bb::R1), bb::mont(3), bb::mont(4), bb::neg(bb::halve(bb::R1)),
        bb::mont_neg(3), bb::mont_neg(4), p2::INV_2_8, bb::halve(bb::halve(bb::R1)), p2::INV_8, p2::INV_2_27,
        bb::neg(p2::INV_2_8), bb::neg(p2::INV_16), bb::neg(p2::INV_2_27)};
    __m256i v0 = _mm256_loadu_si256((const __m256i*)s), v1 = _mm256_loadu_si256((const __m256i*)(s + 8));
    const __m256i d0 = _mm256_loadu_si256((const __m256i*)DIAG), d1 = _mm256_loadu_si256((const __m256i*)(DIAG + 8));
    external_linear(v0, v1);
    for (int r = 0; r < 4; r++) {
        v0 = sbox7(add(v0, _mm256_loadu_si256((const __m256i*)(P2H_EXT_INIT + 16 * r))));
        v1 = sbox7(add(v1, _mm256_loadu_si256((const __m256i*)(P2H_EXT_INIT + 16 * r + 8))));
        external_linear(v0, v1);
    }
    for (int r = 0; r < 13; r++) {
        const uint32_t s0 = scalar_sbox7(bb::add((uint32_t)_mm256_extract_epi32(v0, 0), P2H_INTERNAL[r]));
        v0 = _mm256_insert_epi32(v0, (int)s0, 0);
        __m256i t = add(v0, v1);
        t = add(t, _mm256_permute2x128_si256(t, t, 0x01));
        t = add(t, _mm256_shuffle_epi32(t, 0x4E));
        t = add(t, _mm256_shuffle_epi32(t, 0xB1));  // the sum of the 16 words in every lane
        v0 = add(t, mmul(v0, d0));
        v1 = add(t, mmul(v1, d1));
    }
    for (int r = 0; r < 4; r++) {
        v0 = sbox7(add(v0, _mm256_loadu_si256((const __m256i*)(P2H_EXT_TERM + 16 * r))));
        v1 = sbox7(add(v1, _mm256_loadu_si256((const __m256i*)(P2H_EXT_TERM + 16 * r + 8))));
        external_linear(v0, v1);
    }
    _mm256_storeu_si256((__m256i*)s, v0);
    _mm256_storeu_si256((__m256i*)(s + 8), v1);
}
#endif  // SWIRL_P2_HOST_AVX2

// The permutation the host transcript uses.
static inline void permute(uint32_t s[16]) {
#ifdef SWIRL_P2_HOST_AVX2
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) {
        permute_avx2(s);
        return;
    }
#endif
    p2::permute(s);
}

}  // namespace p2host
