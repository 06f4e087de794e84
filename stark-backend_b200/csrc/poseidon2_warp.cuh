// Poseidon2-BabyBear-16 with the state spread over 16 lanes of a warp (one word per lane): the
// LATENCY-oriented form for a device-side transcript, where one sponge runs alone and the
// thread-per-state permutation (poseidon2_v2.cuh, built for throughput) would leave 31 lanes idle
// while a single thread walks ~5000 dependent-ish instructions (3.5 us on B200).
//
// Here every external round is one S-box per lane (4 chained signed Montgomery products) and five
// shuffles (M4 inside aligned groups of four lanes = group sum + x_k + 2 x_{k+1}; column sums over
// the four groups = two xor-butterflies); an internal round is the S-box on lane 0, overlapped with
// the xor-butterfly sum of the other 15 words, one broadcast and a per-lane diagonal product.
// The critical path is ~190 cycles per external and ~150 per internal round: measured by
// tools/latency_bench.cu.  Bit-exact with p2v2::permute (checked there and by tests).
//
// Replaces (reference, relative to /root/reference): nothing one-to-one; the reference keeps its
// sponge on the host (crates/cuda-backend/src/sponge.rs:267-300) and only grinds on the device.
#pragma once
#include "bb31.cuh"
#include "poseidon2_constants.cuh"

namespace p2w {

#define M(x) bb::mont(x)
static __device__ __constant__ uint32_t W_EXT_INIT[64] = {P2_EXT_INIT_VALUES};
static __device__ __constant__ uint32_t W_INTERNAL[13] = {P2_INTERNAL_VALUES};
static __device__ __constant__ uint32_t W_EXT_TERM[64] = {P2_EXT_TERM_VALUES};
#undef M
// internal diagonal d = (-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 2^-8, 1/4, 1/8, 2^-27, -2^-8, -1/16, -2^-27), Montgomery form
__host__ __device__ constexpr uint32_t inv_pow2(int k) {  // 2^-k mod p, canonical: ((p + 1) / 2)^k
    uint64_t r = 1, h = (uint64_t(bb::P) + 1) / 2;
    for (int i = 0; i < k; i++) r = r * h % bb::P;
    return (uint32_t)r;
}
static __device__ __constant__ uint32_t W_DIAG[16] = {
    bb::mont_neg(2),           bb::mont(1),           bb::mont(2),          bb::mont(inv_pow2(1)),
    bb::mont(3),               bb::mont(4),           bb::mont_neg(inv_pow2(1)), bb::mont_neg(3),
    bb::mont_neg(4),           bb::mont(inv_pow2(8)), bb::mont(inv_pow2(2)), bb::mont(inv_pow2(3)),
    bb::mont(inv_pow2(27)),    bb::mont_neg(inv_pow2(8)), bb::mont_neg(inv_pow2(4)), bb::mont_neg(inv_pow2(27))};

__device__ __forceinline__ uint32_t sbox7(uint32_t x) {  // canonical in, canonical out
    const int32_t s = (int32_t)x;
    const int32_t x2 = bb::smul(s, s), x3 = bb::smul(x2, s), x4 = bb::smul(x2, x2);
    return bb::canon(bb::smul(x3, x4));
}

// `mask` = the 16 participating lanes (an aligned half warp), `lane` = this lane's index in [0, 16)
__device__ __forceinline__ uint32_t external_linear(uint32_t x, unsigned mask) {
    // M4 = circ(2, 3, 1, 1) on each aligned group of four lanes: y_k = S + x_k + 2 x_{k+1}
    const int lane = threadIdx.x & 31;
    const uint32_t nxt = __shfl_sync(mask, x, (lane & ~3) | ((lane + 1) & 3));
    uint32_t s = bb::add(x, __shfl_xor_sync(mask, x, 1));
    s = bb::add(s, __shfl_xor_sync(mask, s, 2));
    const uint32_t y = bb::add(bb::add(s, x), bb::dbl(nxt));
    // out = y + sum over the four groups of the same position
    uint32_t t = bb::add(y, __shfl_xor_sync(mask, y, 4));
    t = bb::add(t, __shfl_xor_sync(mask, t, 8));
    return bb::add(y, t);
}

__device__ __forceinline__ uint32_t permute(uint32_t x, int lane, unsigned mask) {
    x = external_linear(x, mask);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        x = sbox7(bb::add(x, W_EXT_INIT[r * 16 + lane]));
        x = external_linear(x, mask);
    }
    const uint32_t diag = W_DIAG[lane];
    const int base = (threadIdx.x & 31) & ~15;
#pragma unroll 1
    for (int r = 0; r < 13; r++) {
        // lane 0: S-box; meanwhile the other 15 words are summed (lane 0 contributes 0)
        uint32_t rest = lane == 0 ? 0u : x;
        const uint32_t dx = bb::mul(x, diag);  // lanes >= 1: their diagonal term, off the critical path
        if (lane == 0) x = sbox7(bb::add(x, W_INTERNAL[r]));
        rest = bb::add(rest, __shfl_xor_sync(mask, rest, 1));
        rest = bb::add(rest, __shfl_xor_sync(mask, rest, 2));
        rest = bb::add(rest, __shfl_xor_sync(mask, rest, 4));
        rest = bb::add(rest, __shfl_xor_sync(mask, rest, 8));
        const uint32_t y0 = __shfl_sync(mask, x, base);
        // lane 0: -2 y0 + (y0 + rest) = rest - y0;  lane i: d_i x_i + y0 + rest
        x = lane == 0 ? bb::sub(rest, y0) : bb::add(dx, bb::add(rest, y0));
    }
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        x = sbox7(bb::add(x, W_EXT_TERM[r * 16 + lane]));
        x = external_linear(x, mask);
    }
    return x;
}

}  // namespace p2w
