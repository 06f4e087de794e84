// BabyBear (p = 15*2^27 + 1) arithmetic for sm_100a: Montgomery form, radix 2^32, one element per
// 32-bit register, plus the quartic extension EF = F[X]/(X^4 - 11).
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-common/include/fp.h:52-214                 class Fp
//   crates/cuda-common/include/fpext.h:37-121              class FpExt
//   crates/cuda-common/include/ff/baby_bear.hpp:24-400     bb31_t / bb31_4_t (sppark)
// Same in-memory format (word = x * 2^32 mod p, canonical < p), different code: every product
// is one 64-bit IMAD.WIDE, one IMAD for the Montgomery quotient and one IMAD.WIDE that folds
// q*p back in (3 fma-pipe issues), followed by a 2-instruction unsigned-min canonicalisation.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bb {

constexpr uint32_t P = 0x78000001u;
constexpr uint32_t NEG_PINV = 0x77ffffffu;  // -p^-1 mod 2^32
constexpr uint32_t PINV = 0x88000001u;      //  p^-1 mod 2^32
constexpr uint32_t R1 = 0x0ffffffeu;        // 2^32 mod p  (Montgomery 1)
constexpr uint32_t R2 = 1172168163u;        // 2^64 mod p

// compile-time canonical -> Montgomery, for constant tables
__host__ __device__ constexpr uint32_t mont(uint64_t x) { return (uint32_t)(((x % P) << 32) % P); }
__host__ __device__ constexpr uint32_t mont_neg(uint64_t x) { return (x % P) ? P - mont(x) : 0u; }

__host__ __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    uint32_t t = s - P;
    return s < t ? s : t;  // umin
}
__host__ __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) {
    uint32_t d = a - b;
    uint32_t t = d + P;
    return d < t ? d : t;
}
__host__ __device__ __forceinline__ uint32_t neg(uint32_t a) { return a ? P - a : 0u; }
__host__ __device__ __forceinline__ uint32_t dbl(uint32_t a) { return add(a, a); }

// x < p * 2^32  ->  x / 2^32 mod p, canonical
__host__ __device__ __forceinline__ uint32_t reduce(uint64_t x) {
    uint32_t q = (uint32_t)x * NEG_PINV;
    uint64_t t = x + (uint64_t)q * P;
    uint32_t r = (uint32_t)(t >> 32);
    uint32_t u = r - P;
    return r < u ? r : u;
}
// x < 2 * p * 2^32 -> x / 2^32 mod p, canonical.  p*2^32 has a zero low word, so the conditional
// subtraction that brings x below p*2^32 only touches the high word and is one unsigned min there
// (VIADDMNMX) instead of a 64-bit compare + select + subtract (6 instructions).
__host__ __device__ __forceinline__ uint32_t reduce_lazy(uint64_t x) {
    const uint32_t hi = (uint32_t)(x >> 32), hm = hi - P;
    return reduce(((uint64_t)(hi < hm ? hi : hm) << 32) | (uint32_t)x);
}
__host__ __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) { return reduce((uint64_t)a * b); }
__host__ __device__ __forceinline__ uint32_t sqr(uint32_t a) { return mul(a, a); }

// Signed Montgomery product a*b*2^-32 (mod p) for ANY int32 inputs: x = a*b, q = lo(x)*p^-1,
// r = hi(x) - hi(q*p); |r| <= |a||b|/2^32 + p/2, so it is closed on int32 and needs no correction
// between chained products.  On the device x - q*p (a multiple of 2^32) is formed inside the third
// multiply: IMAD.WIDE + IMAD + IMAD.HI with the 64-bit x as addend = 5 fma-heavy passes and NO alu
// op.  ptxas only emits that form when (i) the first product is an explicit mul.wide.s32 (a C++
// int64 product of a value that came out of `>> 32` is expanded into unsigned pieces plus sign
// fix-ups) and (ii) -p is not a visible immediate (knowing q*p == lo(x) it "simplifies" the sum
// into IMAD.HI + a carry chain of three adds), hence the constant-bank operand.
#ifdef __CUDACC__
static __device__ __constant__ int32_t SMUL_NEG_P = -(int32_t)P;
#endif
__host__ __device__ __forceinline__ int32_t smul(int32_t a, int32_t b) {
#ifdef __CUDA_ARCH__
    int64_t x;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(x) : "r"(a), "r"(b));
    const int32_t q = (int32_t)((uint32_t)x * PINV);
    int64_t y;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(y) : "r"(q), "r"(SMUL_NEG_P));
    return (int32_t)((x + y) >> 32);
#else
    const int64_t x = (int64_t)a * b;
    const int32_t q = (int32_t)((uint32_t)x * PINV);
    return (int32_t)(x >> 32) - (int32_t)(((int64_t)q * (int32_t)P) >> 32);
#endif
}
// v in (-p, p) -> canonical [0, p): one VIADDMNMX
__host__ __device__ __forceinline__ uint32_t canon(int32_t v) {
    const uint32_t u = (uint32_t)v, w = u + P;
    return u < w ? u : w;
}
// (u - v) * w for canonical u, v, w: the difference is used unreduced (butterfly lower leg)
__host__ __device__ __forceinline__ uint32_t mul_diff(uint32_t u, uint32_t v, uint32_t w) {
    return canon(smul((int32_t)(u - v), (int32_t)w));
}

__host__ __device__ __forceinline__ uint32_t halve(uint32_t a) {
    // (a + (a odd ? p : 0)) / 2 ; a + p < 2^32
    return (a + ((a & 1u) ? P : 0u)) >> 1;
}

__host__ __device__ __forceinline__ uint32_t to_mont(uint32_t canonical) { return mul(canonical, R2); }
__host__ __device__ __forceinline__ uint32_t from_mont(uint32_t m) { return reduce((uint64_t)m); }

__host__ __device__ inline uint32_t pow(uint32_t b, uint64_t e) {
    uint32_t r = R1;
    while (e) {
        if (e & 1) r = mul(r, b);
        b = sqr(b);
        e >>= 1;
    }
    return r;
}
__host__ __device__ inline uint32_t inv(uint32_t a) { return pow(a, P - 2); }

// Montgomery form of two_adic_generator(bits); canonical generator of the 2^27 subgroup is
// 0x1a427a41 (fp.h:291-320).
__host__ __device__ inline uint32_t two_adic_generator(int bits) {
    uint32_t g = mont(0x1a427a41u);
    for (int i = bits; i < 27; i++) g = sqr(g);
    return g;
}

// ---------------------------------------------------------------------------------------------
// EF = F[X]/(X^4 - 11), basis (1, X, X^2, X^3); 16 bytes, 4-byte aligned in memory.
// ---------------------------------------------------------------------------------------------
struct Ext {
    uint32_t c[4];
};

constexpr uint32_t BETA = mont(11);

__host__ __device__ __forceinline__ Ext ext_zero() { return Ext{{0, 0, 0, 0}}; }
__host__ __device__ __forceinline__ Ext ext_one() { return Ext{{R1, 0, 0, 0}}; }
__host__ __device__ __forceinline__ Ext ext_from(uint32_t a) { return Ext{{a, 0, 0, 0}}; }
__host__ __device__ __forceinline__ Ext ext_add(Ext a, Ext b) {
    return Ext{{add(a.c[0], b.c[0]), add(a.c[1], b.c[1]), add(a.c[2], b.c[2]), add(a.c[3], b.c[3])}};
}
__host__ __device__ __forceinline__ Ext ext_sub(Ext a, Ext b) {
    return Ext{{sub(a.c[0], b.c[0]), sub(a.c[1], b.c[1]), sub(a.c[2], b.c[2]), sub(a.c[3], b.c[3])}};
}
__host__ __device__ __forceinline__ Ext ext_neg(Ext a) {
    return Ext{{neg(a.c[0]), neg(a.c[1]), neg(a.c[2]), neg(a.c[3])}};
}
__host__ __device__ __forceinline__ Ext ext_mul_base(Ext a, uint32_t s) {
    return Ext{{mul(a.c[0], s), mul(a.c[1], s), mul(a.c[2], s), mul(a.c[3], s)}};
}

// Sum of four 62-bit products with a single Montgomery reduction.  Each a_i*b_i < p^2, and
// 4 p^2 < 2^64, so the plain 64-bit sum cannot wrap; one conditional subtraction of p*2^32
// (which changes neither the residue mod p nor the low word) brings it below p*2^32 as
// `reduce` requires (4 p^2 - p 2^32 < p 2^32).
__host__ __device__ __forceinline__ uint32_t dot4(uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1,
                                                  uint32_t a2, uint32_t b2, uint32_t a3, uint32_t b3) {
    const uint64_t s = ((uint64_t)a0 * b0 + (uint64_t)a1 * b1) + ((uint64_t)a2 * b2 + (uint64_t)a3 * b3);
    return reduce_lazy(s);
}

__host__ __device__ __forceinline__ Ext ext_mul(Ext a, Ext b) {
    // c0 = a0b0 + 11(a1b3 + a2b2 + a3b1)
    // c1 = a0b1 + a1b0 + 11(a2b3 + a3b2)
    // c2 = a0b2 + a1b1 + a2b0 + 11 a3b3
    // c3 = a0b3 + a1b2 + a2b1 + a3b0
    uint32_t w1 = mul(b.c[1], BETA), w2 = mul(b.c[2], BETA), w3 = mul(b.c[3], BETA);
    Ext r;
    r.c[0] = dot4(a.c[0], b.c[0], a.c[1], w3, a.c[2], w2, a.c[3], w1);
    r.c[1] = dot4(a.c[0], b.c[1], a.c[1], b.c[0], a.c[2], w3, a.c[3], w2);
    r.c[2] = dot4(a.c[0], b.c[2], a.c[1], b.c[1], a.c[2], b.c[0], a.c[3], w3);
    r.c[3] = dot4(a.c[0], b.c[3], a.c[1], b.c[2], a.c[2], b.c[1], a.c[3], b.c[0]);
    return r;
}
__host__ __device__ __forceinline__ Ext ext_sqr(Ext a) { return ext_mul(a, a); }

__host__ __device__ inline Ext ext_inv(Ext a) {
    // norm to F[Y]/(Y^2-11), Y = X^2:  a = A + X B,  a (A - X B) = A^2 - Y B^2
    uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    uint32_t A0 = add(sqr(a0), mul(BETA, sqr(a2))), A1 = dbl(mul(a0, a2));
    uint32_t B0 = add(sqr(a1), mul(BETA, sqr(a3))), B1 = dbl(mul(a1, a3));
    uint32_t n0 = sub(A0, mul(BETA, B1)), n1 = sub(A1, B0);
    uint32_t d = inv(sub(sqr(n0), mul(BETA, sqr(n1))));
    uint32_t i0 = mul(n0, d), i1 = neg(mul(n1, d));
    Ext r;
    r.c[0] = add(mul(a0, i0), mul(BETA, mul(a2, i1)));
    r.c[2] = add(mul(a0, i1), mul(a2, i0));
    r.c[1] = neg(add(mul(a1, i0), mul(BETA, mul(a3, i1))));
    r.c[3] = neg(add(mul(a1, i1), mul(a3, i0)));
    return r;
}

}  // namespace bb
