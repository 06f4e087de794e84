// Poseidon2 over BabyBear, width 16, x^7, 4 + 13 + 4 rounds — register-resident permutation for
// sm_100a, and the two hash modes of the SWIRL Merkle commitment.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-common/include/poseidon2.cuh:77-202         poseidon2::poseidon2_mix
//   crates/cuda-backend/cuda/src/merkle_tree.cu:32-51        leaf sponge
//   crates/cuda-backend/cuda/src/merkle_tree.cu:166-172      2-to-1 compress
// The whole state lives in 16 registers; round constants sit in the constant bank and are read
// as instruction operands (warp-uniform addresses).  `__host__ __device__` so that the host-side
// transcript and host unit tests run exactly this code.
#pragma once
#include "bb31.cuh"
#include "poseidon2_constants.cuh"

namespace p2 {

#ifdef __CUDA_ARCH__
#define P2_RC_INIT(i) P2C_EXT_INIT[i]
#define P2_RC_INT(i) P2C_INTERNAL[i]
#define P2_RC_TERM(i) P2C_EXT_TERM[i]
#else
#define P2_RC_INIT(i) P2H_EXT_INIT[i]
#define P2_RC_INT(i) P2H_INTERNAL[i]
#define P2_RC_TERM(i) P2H_EXT_TERM[i]
#endif

// Montgomery words of the non-trivial internal-diagonal entries
constexpr uint32_t INV_2_8 = bb::mont(2005401601u);    // 2^-8
constexpr uint32_t INV_8 = bb::mont(1761607681u);      // 1/8
constexpr uint32_t INV_2_27 = bb::mont(2013265906u);   // 2^-27
constexpr uint32_t INV_16 = bb::mont(1887436801u);     // 1/16

__host__ __device__ __forceinline__ uint32_t sbox7(uint32_t x) {
    uint32_t x2 = bb::sqr(x);
    uint32_t x3 = bb::mul(x2, x);
    uint32_t x4 = bb::sqr(x2);
    return bb::mul(x3, x4);
}

// y = M4 x with M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]], 4 rows sharing partial sums
__host__ __device__ __forceinline__ void m4(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    uint32_t ab = bb::add(a, b), cd = bb::add(c, d);
    uint32_t all = bb::add(ab, cd);
    uint32_t ya = bb::add(bb::add(all, ab), b);           // 2a+3b+c+d
    uint32_t yb = bb::add(bb::add(all, b), bb::dbl(c));   // a+2b+3c+d
    uint32_t yc = bb::add(bb::add(all, cd), d);           // a+b+2c+3d
    uint32_t yd = bb::add(bb::add(all, d), bb::dbl(a));   // 3a+b+c+2d
    a = ya;
    b = yb;
    c = yc;
    d = yd;
}

__host__ __device__ __forceinline__ void external_linear(uint32_t s[16]) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) m4(s[i], s[i + 1], s[i + 2], s[i + 3]);
    uint32_t t0 = bb::add(bb::add(s[0], s[4]), bb::add(s[8], s[12]));
    uint32_t t1 = bb::add(bb::add(s[1], s[5]), bb::add(s[9], s[13]));
    uint32_t t2 = bb::add(bb::add(s[2], s[6]), bb::add(s[10], s[14]));
    uint32_t t3 = bb::add(bb::add(s[3], s[7]), bb::add(s[11], s[15]));
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        s[i] = bb::add(s[i], t0);
        s[i + 1] = bb::add(s[i + 1], t1);
        s[i + 2] = bb::add(s[i + 2], t2);
        s[i + 3] = bb::add(s[i + 3], t3);
    }
}

// s <- (J + diag(d)) s,  d = (-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 2^-8, 1/4, 1/8, 2^-27, -2^-8, -1/16, -2^-27)
__host__ __device__ __forceinline__ void internal_linear(uint32_t s[16]) {
    uint32_t rest = bb::add(bb::add(bb::add(s[1], s[2]), bb::add(s[3], s[4])),
                            bb::add(bb::add(s[5], s[6]), bb::add(s[7], s[8])));
    uint32_t rest2 = bb::add(bb::add(bb::add(s[9], s[10]), bb::add(s[11], s[12])),
                             bb::add(bb::add(s[13], s[14]), s[15]));
    rest = bb::add(rest, rest2);
    uint32_t sum = bb::add(rest, s[0]);
    s[0] = bb::sub(rest, s[0]);
    s[1] = bb::add(sum, s[1]);
    s[2] = bb::add(sum, bb::dbl(s[2]));
    s[3] = bb::add(sum, bb::halve(s[3]));
    s[4] = bb::add(sum, bb::add(bb::dbl(s[4]), s[4]));
    s[5] = bb::add(sum, bb::dbl(bb::dbl(s[5])));
    s[6] = bb::sub(sum, bb::halve(s[6]));
    s[7] = bb::sub(sum, bb::add(bb::dbl(s[7]), s[7]));
    s[8] = bb::sub(sum, bb::dbl(bb::dbl(s[8])));
    s[9] = bb::add(sum, bb::mul(s[9], INV_2_8));
    s[10] = bb::add(sum, bb::halve(bb::halve(s[10])));
    s[11] = bb::add(sum, bb::mul(s[11], INV_8));
    s[12] = bb::add(sum, bb::mul(s[12], INV_2_27));
    s[13] = bb::sub(sum, bb::mul(s[13], INV_2_8));
    s[14] = bb::sub(sum, bb::mul(s[14], INV_16));
    s[15] = bb::sub(sum, bb::mul(s[15], INV_2_27));
}

// First-generation permutation (every value canonical, unsigned Montgomery products); kept as the
// in-library cross-check of p2v2::permute (tools/p2_bench.cu compares them on device).
__host__ __device__ __forceinline__ void permute_v1(uint32_t s[16]) {
    external_linear(s);
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox7(bb::add(s[i], P2_RC_INIT(r * 16 + i)));
        external_linear(s);
    }
#pragma unroll
    for (int r = 0; r < 13; r++) {
        s[0] = sbox7(bb::add(s[0], P2_RC_INT(r)));
        internal_linear(s);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox7(bb::add(s[i], P2_RC_TERM(r * 16 + i)));
        external_linear(s);
    }
}

}  // namespace p2

#include "poseidon2_v2.cuh"

namespace p2 {
// The permutation every kernel and the host transcript use (3.85 vs 2.89 Gperm/s for v1 on B200).
__host__ __device__ __forceinline__ void permute(uint32_t s[16]) { p2v2::permute(s); }
}  // namespace p2
