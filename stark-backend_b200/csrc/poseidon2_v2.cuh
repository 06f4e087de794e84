// Poseidon2 over BabyBear, width 16, x^7, 4 + 13 + 4 rounds — second-generation permutation for
// sm_100a, tuned against the measured integer issue rates (tools/int_roofline.cu):
//   IMAD 1 fma-pipe slot, IMAD.WIDE 2, IMAD.HI 2, IADD3/LOP3/SHF/VIADDMNMX 1 alu-pipe slot,
//   64 lanes/clk/SM per pipe.  A canonical Montgomery product is 5 fma slots + 2 alu ops; the
//   permutation has 564 S-box products, so the fma pipe is the floor and everything else is
//   arranged to stay off it and to use as few alu ops as possible:
//   * S-box in *signed* Montgomery form (x = a*b; q = lo(x)*p^-1; r = hi(x) - hi(q*p)): closed on
//     int32 with |r| <= |a||b|/2^32 + p/2, so the three intermediate powers need no correction and
//     only x^7 is brought back to [0,p) with one VIADDMNMX;
//   * the round constant is stored as rc - p in (-p, 0], which makes `state + rc` a single add
//     with a result in [-p, p), and for the external layers the last add of the linear layer and
//     the next round's constant are one 3-input IADD3;
//   * multiplications by 2^-k in the internal diagonal use p = 15*2^27 + 1:
//     x*2^-k = (x >> k) - 15*2^(27-k) * (x mod 2^k), one shift, one mask, one IMAD;
//   * round loops stay rolled (`#pragma unroll 1`) so the body is ~20 KB of code and lives in
//     the instruction cache instead of streaming 100+ KB of straight-line SASS per warp.
// Replaces (reference, relative to /root/reference):
//   crates/cuda-common/include/poseidon2.cuh:77-202         poseidon2::poseidon2_mix
// Bit-exact with p2::permute (poseidon2.cuh) and the CPU oracle; canonical Montgomery words in
// and out.
#pragma once
#include "bb31.cuh"
#include "poseidon2_constants.cuh"

namespace p2v2 {

constexpr int32_t P = (int32_t)bb::P;

// rc - p as signed words (rc Montgomery canonical)
#define M(x) ((int32_t)bb::mont(x) - (int32_t)bb::P)
// a fifth block of "zero" constants (0 - p) lets the last external layer of each half run the same
// fused code; its outputs are canonicalised afterwards
#define Z4 -P, -P, -P, -P
#define Z16 Z4, Z4, Z4, Z4
static __device__ __constant__ int32_t C_EXT_INIT[80] = {P2_EXT_INIT_VALUES Z16};
static __device__ __constant__ int32_t C_INTERNAL[13] = {P2_INTERNAL_VALUES};
static __device__ __constant__ int32_t C_EXT_TERM[80] = {P2_EXT_TERM_VALUES Z16};
static const int32_t H_EXT_INIT[80] = {P2_EXT_INIT_VALUES Z16};
static const int32_t H_INTERNAL[13] = {P2_INTERNAL_VALUES};
static const int32_t H_EXT_TERM[80] = {P2_EXT_TERM_VALUES Z16};
#undef Z16
#undef Z4
#undef M

#ifdef __CUDA_ARCH__
#define P2V2_RC_INIT(i) C_EXT_INIT[i]
#define P2V2_RC_INT(i) C_INTERNAL[i]
#define P2V2_RC_TERM(i) C_EXT_TERM[i]
#else
#define P2V2_RC_INIT(i) H_EXT_INIT[i]
#define P2V2_RC_INT(i) H_INTERNAL[i]
#define P2V2_RC_TERM(i) H_EXT_TERM[i]
#endif

using bb::canon;
using bb::smul;

// ---- pipe placement of the plain additions ----------------------------------------------------------
// ptxas splits two-input adds between the ALU pipe (IADD3) and the fma-heavy pipe (IMAD.IADD) by
// instruction COUNT; it does not know that IMAD.WIDE / IMAD.HI occupy the heavy pipe for two passes,
// so its split leaves the multiplier pipe ~25 % busier than the ALU pipe.  An add with a third
// operand that is zero at run time (constant bank) can only be an IADD3, which pins it to the ALU.
// POL is a bit mask of the groups of additions that are pinned (tools/p2_bench.cu sweeps it).
constexpr int PIN_M4 = 1, PIN_COLSUM = 2, PIN_EXT_OUT = 4, PIN_INT_SUM = 8, PIN_INT_OUT = 16, PIN_EXT_RC = 32;
#ifndef P2V2_PIN_POLICY
#define P2V2_PIN_POLICY 0
#endif
#ifdef __CUDACC__
static __device__ __constant__ uint32_t PIN_ZERO = 0;
#endif
template <bool PIN>
__host__ __device__ __forceinline__ uint32_t radd(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    if (PIN) return a + b + PIN_ZERO;
#endif
    return a + b;
}
template <bool PIN>
__host__ __device__ __forceinline__ uint32_t rsub(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    if (PIN) return a - b + PIN_ZERO;
#endif
    return a - b;
}
template <bool PIN>
__host__ __device__ __forceinline__ uint32_t padd(uint32_t a, uint32_t b) {
    const uint32_t s = radd<PIN>(a, b), t = s - bb::P;
    return s < t ? s : t;
}
template <bool PIN>
__host__ __device__ __forceinline__ uint32_t psub(uint32_t a, uint32_t b) {
    const uint32_t d = rsub<PIN>(a, b), t = d + bb::P;
    return d < t ? d : t;
}
template <bool PIN>
__host__ __device__ __forceinline__ uint32_t pdbl(uint32_t a) { return padd<PIN>(a, a); }

// s in [-p, p) (state word plus signed round constant) -> s^7 canonical
__host__ __device__ __forceinline__ uint32_t sbox7(int32_t s) {
    const int32_t x2 = smul(s, s);    // |.| < 0.97 p
    const int32_t x3 = smul(x2, s);   // < 0.96 p
    const int32_t x4 = smul(x2, x2);  // < 0.95 p
    return canon(smul(x3, x4));       // < 0.93 p
}

// y = M4 x, M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]]; 11 canonical additions
template <int POL>
__host__ __device__ __forceinline__ void m4(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3) {
    constexpr bool Q = (POL & PIN_M4) != 0;
    const uint32_t t01 = padd<Q>(x0, x1), t23 = padd<Q>(x2, x3);
    const uint32_t t0123 = padd<Q>(t01, t23);
    const uint32_t t01123 = padd<Q>(t0123, x1), t01233 = padd<Q>(t0123, x3);
    const uint32_t y3 = padd<Q>(t01233, pdbl<Q>(x0));
    const uint32_t y1 = padd<Q>(t01123, pdbl<Q>(x2));
    x0 = padd<Q>(t01123, t01);
    x2 = padd<Q>(t01233, t23);
    x1 = y1;
    x3 = y3;
}

// External linear layer fused with the next round's constants: the outputs are
// state + (rc - p) in [-p, p), ready for the S-box (one IADD3 per word).
template <int POL>
__host__ __device__ __forceinline__ void external_linear_rc(uint32_t s[16], const int32_t* rc) {
    constexpr bool QC = (POL & PIN_COLSUM) != 0, QO = (POL & PIN_EXT_OUT) != 0;
#pragma unroll
    for (int i = 0; i < 16; i += 4) m4<POL>(s[i], s[i + 1], s[i + 2], s[i + 3]);
    uint32_t t[4];
#pragma unroll
    for (int j = 0; j < 4; j++) t[j] = padd<QC>(padd<QC>(s[j], s[4 + j]), padd<QC>(s[8 + j], s[12 + j]));
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = radd<(POL & PIN_EXT_RC) != 0>(padd<QO>(s[i], t[i & 3]), (uint32_t)rc[i]);  // [0,p) + [-p,0)
}

// x * 2^-k for canonical x, 1 <= k <= 27:  (x >> k) - 15 * 2^(27-k) * (x mod 2^k), in (-p, 2^30]
template <int K>
__host__ __device__ __forceinline__ int32_t mul_2exp_neg_signed(uint32_t x) {
    const uint32_t hi = x >> K, lo = x & ((1u << K) - 1);
    return (int32_t)(hi - lo * (15u << (27 - K)));
}
template <int K>
__host__ __device__ __forceinline__ uint32_t mul_2exp_neg(uint32_t x) {
    return canon(mul_2exp_neg_signed<K>(x));
}
// -(x * 2^-k), canonical
template <int K>
__host__ __device__ __forceinline__ uint32_t mul_neg_2exp_neg(uint32_t x) {
    const uint32_t hi = x >> K, lo = x & ((1u << K) - 1);
    return canon((int32_t)(lo * (15u << (27 - K)) - hi));  // in [-2^30, p)
}

// s <- (J + diag(d)) s, d = (-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 2^-8, 1/4, 1/8, 2^-27, -2^-8, -1/16, -2^-27)
template <int POL>
__host__ __device__ __forceinline__ void internal_linear(uint32_t s[16]) {
    constexpr bool QS = (POL & PIN_INT_SUM) != 0, Q = (POL & PIN_INT_OUT) != 0;
    const uint32_t r1 = padd<QS>(padd<QS>(padd<QS>(s[1], s[2]), padd<QS>(s[3], s[4])), padd<QS>(padd<QS>(s[5], s[6]), padd<QS>(s[7], s[8])));
    const uint32_t r2 = padd<QS>(padd<QS>(padd<QS>(s[9], s[10]), padd<QS>(s[11], s[12])), padd<QS>(padd<QS>(s[13], s[14]), s[15]));
    const uint32_t rest = padd<QS>(r1, r2);
    const uint32_t sum = padd<QS>(rest, s[0]);
    s[0] = psub<Q>(rest, s[0]);
    s[1] = padd<Q>(sum, s[1]);
    s[2] = padd<Q>(padd<Q>(sum, s[2]), s[2]);
    s[3] = padd<Q>(sum, bb::halve(s[3]));
    s[4] = padd<Q>(padd<Q>(sum, s[4]), pdbl<Q>(s[4]));
    {
        const uint32_t d = pdbl<Q>(s[5]);
        s[5] = padd<Q>(padd<Q>(sum, d), d);
    }
    s[6] = psub<Q>(sum, bb::halve(s[6]));
    s[7] = psub<Q>(psub<Q>(sum, s[7]), pdbl<Q>(s[7]));
    {
        const uint32_t d = pdbl<Q>(s[8]);
        s[8] = psub<Q>(psub<Q>(sum, d), d);
    }
    s[9] = padd<Q>(sum, mul_2exp_neg<8>(s[9]));
    s[10] = padd<Q>(sum, mul_2exp_neg<2>(s[10]));
    s[11] = padd<Q>(sum, mul_2exp_neg<3>(s[11]));
    s[12] = padd<Q>(sum, mul_2exp_neg<27>(s[12]));
    s[13] = padd<Q>(sum, mul_neg_2exp_neg<8>(s[13]));
    s[14] = padd<Q>(sum, mul_neg_2exp_neg<4>(s[14]));
    s[15] = padd<Q>(sum, mul_neg_2exp_neg<27>(s[15]));
}

#ifdef P2V2_UNROLL_ROUNDS
#define P2V2_ROUND_LOOP _Pragma("unroll")
#else
#define P2V2_ROUND_LOOP _Pragma("unroll 1")
#endif

template <int POL>
__host__ __device__ __forceinline__ void permute_pol(uint32_t s[16]) {
    external_linear_rc<POL>(s, &P2V2_RC_INIT(0));
    P2V2_ROUND_LOOP
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox7((int32_t)s[i]);
        external_linear_rc<POL>(s, &P2V2_RC_INIT((r + 1) * 16));
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = canon((int32_t)s[i]);
    P2V2_ROUND_LOOP
    for (int r = 0; r < 13; r++) {
        s[0] = sbox7((int32_t)s[0] + P2V2_RC_INT(r));
        internal_linear<POL>(s);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = (uint32_t)((int32_t)s[i] + P2V2_RC_TERM(i));
    P2V2_ROUND_LOOP
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox7((int32_t)s[i]);
        external_linear_rc<POL>(s, &P2V2_RC_TERM((r + 1) * 16));
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = canon((int32_t)s[i]);
}

__host__ __device__ __forceinline__ void permute(uint32_t s[16]) { permute_pol<P2V2_PIN_POLICY>(s); }

}  // namespace p2v2
