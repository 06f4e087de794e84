// Poseidon2 over BabyBear (width 16, x^7, 4 + 13 + 4 rounds) on the FP64 pipe of sm_100a.
//
// Why: the integer permutation (poseidon2_v2.cuh) is bound by the fma-heavy pipe (32-bit integer multiplies,
// 64 lanes/clk/SM) with the alu pipe as co-limit; B200 also has a full-rate FP64 pipe (64 DFMA lanes/clk/SM, ~37
// TFLOP/s) that integer code leaves idle.  This file is the same permutation written for that pipe, so that FP64
// warps can hash rows next to integer warps on the same SM (merkle.cu: leaf_tree_kernel hands rows out dynamically).
//
// Representation: a field element is a double holding an *integer* congruent to the value mod p (standard, not
// Montgomery, domain), lazily reduced: every double stays far below 2^53 in magnitude, so additions, subtractions
// and multiplications by the small integers of the linear layers are exact integer arithmetic.
//   mulmod(a, b)   h = a*b (rounded), l = fma(a, b, -h) (exact low part), q = rint(h / p) via the 1.5*2^52 magic
//                  constant, r = fma(-q, p, h) + l.  h - q*p is an integer below 2^31 in magnitude, hence exact; the
//                  result is congruent to a*b with |r| <= p/2 + |l| + eps.  6 FP64 issues, valid while |a*b| < 2^80.
//   x * 2^-k       p = 15 * 2^27 + 1:  x = xr * 2^k + lo  =>  x * 2^-k = xr - 15 * 2^(27-k) * lo  (4 issues, no
//                  product reduction), |result| <= |x| / 2^k + p/2.
// The output is bit-identical to p2v2::permute / the CPU oracle / the reference's poseidon2_mix (integer arithmetic
// on exact integers; tests/test_gpu_commit_path.py, tools/p2_fp64_bench.cu).
// Replaces (reference, relative to /root/reference): crates/cuda-common/include/poseidon2.cuh:77-202.
#pragma once
#include "bb31.cuh"
#include "poseidon2_constants.cuh"

namespace p2f {

constexpr double PD = 2013265921.0;            // p
constexpr double PINV = 1.0 / 2013265921.0;    // rounded 1/p
constexpr double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: (x + MAGIC) - MAGIC = rint(x) for |x| < 2^51
constexpr double RINV = 943718400.0;           // 2^-32 mod p  (Montgomery word -> standard value)
constexpr double RMOD = 268435454.0;           // 2^32 mod p   (standard value -> Montgomery word)
static_assert((uint64_t)943718400u * ((1ull << 32) % 2013265921ull) % 2013265921ull == 1, "RINV");

// canonical constants, moved to the balanced range (-p/2, p/2] to keep magnitudes small
#define M(x) ((x) > 1006632960u ? (double)(x) - 2013265921.0 : (double)(x))
static __device__ __constant__ double D_EXT_INIT[64] = {P2_EXT_INIT_VALUES};
static __device__ __constant__ double D_INTERNAL[13] = {P2_INTERNAL_VALUES};
static __device__ __constant__ double D_EXT_TERM[64] = {P2_EXT_TERM_VALUES};
#undef M

__device__ __forceinline__ double rint_magic(double x_times_scale_plus_magic) {
    return __dadd_rn(x_times_scale_plus_magic, -MAGIC);
}

// a * b mod p (see header); 6 issues
__device__ __forceinline__ double mulmod(double a, double b) {
    const double h = __dmul_rn(a, b);
    const double l = __fma_rn(a, b, -h);
    const double q = rint_magic(__fma_rn(h, PINV, MAGIC));
    return __dadd_rn(__fma_rn(-q, PD, h), l);
}
// x mod p into (-p/2 - 1, p/2 + 1); 3 issues, valid for |x| < 2^80
__device__ __forceinline__ double reduce(double x) {
    const double q = rint_magic(__fma_rn(x, PINV, MAGIC));
    return __fma_rn(-q, PD, x);
}
__device__ __forceinline__ double sbox7(double x) {
    const double x2 = mulmod(x, x);
    const double x3 = mulmod(x2, x);
    const double x4 = mulmod(x2, x2);
    return mulmod(x3, x4);
}

// +- x * 2^-K mod p, 1 <= K <= 27
template <int K, bool NEG>
__device__ __forceinline__ double mul_2exp_neg(double x) {
    constexpr double INV2K = 1.0 / (double)(1u << K), TWO_K = (double)(1u << K), C = 15.0 * (double)(1u << (27 - K));
    const double xr = rint_magic(__fma_rn(x, INV2K, MAGIC));
    const double lo = __fma_rn(xr, -TWO_K, x);
    return NEG ? __fma_rn(lo, C, -xr) : __fma_rn(lo, -C, xr);
}

// y = M4 x, M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]]: 9 issues (exact integer arithmetic)
__device__ __forceinline__ void m4(double& x0, double& x1, double& x2, double& x3) {
    const double t01 = __dadd_rn(x0, x1), t23 = __dadd_rn(x2, x3);
    const double t0123 = __dadd_rn(t01, t23);
    const double t01123 = __dadd_rn(t0123, x1), t01233 = __dadd_rn(t0123, x3);
    const double y3 = __fma_rn(x0, 2.0, t01233);
    const double y1 = __fma_rn(x2, 2.0, t01123);
    x0 = __dadd_rn(t01123, t01);
    x2 = __dadd_rn(t01233, t23);
    x1 = y1;
    x3 = y3;
}
// external linear layer: |out| <= 35 * max|in|
__device__ __forceinline__ void external_linear(double s[16]) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) m4(s[i], s[i + 1], s[i + 2], s[i + 3]);
    double t[4];
#pragma unroll
    for (int j = 0; j < 4; j++) t[j] = __dadd_rn(__dadd_rn(s[j], s[4 + j]), __dadd_rn(s[8 + j], s[12 + j]));
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = __dadd_rn(s[i], t[i & 3]);
}

// s <- (J + diag(d)) s, d = (-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 2^-8, 1/4, 1/8, 2^-27, -2^-8, -1/16, -2^-27).
// The sum is reduced (so it adds < p/2 + 1 to every word); integer multiples grow a word at most 4x per round.
__device__ __forceinline__ void internal_linear(double s[16]) {
    const double r1 = __dadd_rn(__dadd_rn(__dadd_rn(s[0], s[1]), __dadd_rn(s[2], s[3])),
                                __dadd_rn(__dadd_rn(s[4], s[5]), __dadd_rn(s[6], s[7])));
    const double r2 = __dadd_rn(__dadd_rn(__dadd_rn(s[8], s[9]), __dadd_rn(s[10], s[11])),
                                __dadd_rn(__dadd_rn(s[12], s[13]), __dadd_rn(s[14], s[15])));
    const double sum = reduce(__dadd_rn(r1, r2));
    s[0] = __fma_rn(s[0], -2.0, sum);
    s[1] = __dadd_rn(s[1], sum);
    s[2] = __fma_rn(s[2], 2.0, sum);
    s[3] = __dadd_rn(mul_2exp_neg<1, false>(s[3]), sum);
    s[4] = __fma_rn(s[4], 3.0, sum);
    s[5] = __fma_rn(s[5], 4.0, sum);
    s[6] = __dadd_rn(mul_2exp_neg<1, true>(s[6]), sum);
    s[7] = __fma_rn(s[7], -3.0, sum);
    s[8] = __fma_rn(s[8], -4.0, sum);
    s[9] = __dadd_rn(mul_2exp_neg<8, false>(s[9]), sum);
    s[10] = __dadd_rn(mul_2exp_neg<2, false>(s[10]), sum);
    s[11] = __dadd_rn(mul_2exp_neg<3, false>(s[11]), sum);
    s[12] = __dadd_rn(mul_2exp_neg<27, false>(s[12]), sum);
    s[13] = __dadd_rn(mul_2exp_neg<8, true>(s[13]), sum);
    s[14] = __dadd_rn(mul_2exp_neg<4, true>(s[14]), sum);
    s[15] = __dadd_rn(mul_2exp_neg<27, true>(s[15]), sum);
}

#ifdef P2F_UNROLL_ROUNDS
#define P2F_ROUND_LOOP _Pragma("unroll")
#else
#define P2F_ROUND_LOOP _Pragma("unroll 1")
#endif

// In: |s[i]| < 2^32 (e.g. canonical values or the previous output).  Out: |s[i]| < 2^36, congruent to the
// permutation output; `finish()` brings words to canonical Montgomery form.
__device__ __forceinline__ void permute(double s[16]) {
    external_linear(s);  // < 2^38
    P2F_ROUND_LOOP
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox7(__dadd_rn(s[i], D_EXT_INIT[r * 16 + i]));  // |.| < p/2 + 2^20
        external_linear(s);                                                                // < 2^35.2
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = reduce(s[i]);
    P2F_ROUND_LOOP
    for (int r = 0; r < 13; r++) {
        s[0] = sbox7(__dadd_rn(s[0], D_INTERNAL[r]));
        internal_linear(s);  // words with |d| = 4 grow 2 bits per round: 2^30 -> 2^44 after 7 rounds
        if (r == 6) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = reduce(s[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = reduce(s[i]);
    P2F_ROUND_LOOP
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox7(__dadd_rn(s[i], D_EXT_TERM[r * 16 + i]));
        external_linear(s);
    }
}

// canonical Montgomery word (what the integer code and HBM hold) -> lazily reduced standard value
__device__ __forceinline__ double from_word(uint32_t m) { return mulmod((double)m, RINV); }
// lazily reduced standard value -> canonical Montgomery word
__device__ __forceinline__ uint32_t to_word(double x) {
    double r = mulmod(x, RMOD);  // (-p/2 - 2^8, p/2 + 2^8)
    r = r < 0.0 ? __dadd_rn(r, PD) : r;
    return (uint32_t)__double2uint_rn(r);
}

}  // namespace p2f
