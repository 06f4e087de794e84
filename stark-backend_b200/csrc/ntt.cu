// Batched number-theoretic transforms over BabyBear for column-major matrices, and the
// Reed–Solomon encoder of the stacked PCS built on them.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-backend/cuda/src/batch_ntt_small.cu:78-153    batch_ntt_kernel (2^l_skip chunks)
//   crates/cuda-backend/cuda/supra/ntt.cu:22-254              _CT_NTT mixed radix passes (sppark)
//   crates/cuda-backend/cuda/supra/ntt_bitrev.cu:47-222       bit_rev_permutation
//   crates/cuda-backend/cuda/src/mle_interpolate.cu:16-445    zeta stages inside chunks
//   crates/cuda-backend/cuda/src/matrix.cu (batch_expand_pad) zero padding to codeword height
//   crates/cuda-backend/src/stacked_pcs.rs:229-337            rs_code_matrix orchestration
//   crates/cuda-backend/src/ntt.rs:111-168                    pass planning
// Semantics = prover/stacked_pcs.rs:341-367 + prover/poly.rs:325-348 (natural order in and out,
// out[i] = sum_j c_j w^(ij), w = two_adic_generator(log n)).
//
// Design (not the reference's): a size-N transform is split into at most three passes
// N = R1*R2*R3 (four-step / six-step style, decimation in frequency).  A pass loads a
// [R x TW] tile (TW contiguous elements per row => coalesced), transforms along R in shared
// memory, multiplies by the inter-pass twiddle and stores.  The last pass reads whole contiguous
// rows and writes the digit-reversed (natural) positions TW at a time, so no separate bit-reversal
// or transpose sweep exists.  Zero padding of the RS message is never materialised: the first pass
// simply treats rows beyond the message as zero.  Columns are processed in groups whose scratch
// fits in L2, so the intermediate between passes does not travel to HBM.
#include <algorithm>
#include <type_traits>
#include <vector>

#include "bb31.cuh"
#include "kernels.cuh"

namespace swirl {

constexpr int NTT_THREADS = 512;
constexpr int NTT_TILE_ELEMS = 16384;  // elements of shared memory per tile (before padding)
constexpr uint32_t W27_MASK = (1u << 27) - 1;
constexpr uint32_t TW_HI_MASK = (1u << TW_HI_BITS) - 1;

// w^E for the 2^27-th root w, E < 2^27, via two tables.  `lo` may be a pre-scaled copy of the low
// table (lo[i] = c * w^i), which folds a constant factor into the same single multiplication.
template <bool SCALED>
__device__ __forceinline__ uint32_t root_pow(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi,
                                             uint32_t E) {
    const uint32_t h = __ldg(hi + (E >> TW_LO_BITS));
    const uint32_t l = E & ((1u << TW_LO_BITS) - 1);
    if (SCALED) return bb::mul(h, __ldg(lo + l));
    return l ? bb::mul(h, __ldg(lo + l)) : h;
}

__global__ void twiddle_init_kernel(uint32_t* lo, uint32_t* hi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = bb::two_adic_generator(27);
    if (i < (1u << TW_LO_BITS)) lo[i] = bb::pow(w, i);
    if (i < (1u << TW_HI_BITS)) hi[i] = bb::pow(w, (uint64_t)i << TW_LO_BITS);
}
__global__ void twiddle_scale_kernel(const uint32_t* __restrict__ lo, uint32_t* __restrict__ out, uint32_t c) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (1u << TW_LO_BITS)) out[i] = bb::mul(lo[i], c);
}

int ntt_init_twiddles(swirl_ctx* ctx) {
    SWIRL_CUDA(cudaMalloc((void**)&ctx->tw_lo, sizeof(uint32_t) << TW_LO_BITS));
    SWIRL_CUDA(cudaMalloc((void**)&ctx->tw_hi, sizeof(uint32_t) << TW_HI_BITS));
    twiddle_init_kernel<<<(1u << TW_LO_BITS) / 256, 256, 0, ctx->stream>>>(ctx->tw_lo, ctx->tw_hi);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}

// tw_lo scaled by 2^-l (the 1/|D| of the chunk iDFT), built on first use
static int scaled_twiddles(swirl_ctx* ctx, int l, const uint32_t** out) {
    if (!ctx->tw_lo_scaled[l]) {
        uint32_t* t = nullptr;
        SWIRL_CUDA(cudaMalloc((void**)&t, sizeof(uint32_t) << TW_LO_BITS));
        twiddle_scale_kernel<<<(1u << TW_LO_BITS) / 256, 256, 0, ctx->stream>>>(ctx->tw_lo, t,
                                                                              bb::inv(bb::to_mont(1u << l)));
        ctx->tw_lo_scaled[l] = t;
        SWIRL_LAUNCH_CHECK(ctx);
    }
    *out = ctx->tw_lo_scaled[l];
    return 0;
}

struct PassArgs {
    const uint32_t* src;
    uint32_t* dst;
    size_t src_col_stride, dst_col_stride;
    uint32_t cols;
    int log_n;         // whole transform
    int log_r;         // radix of this pass
    int log_s;         // element stride of the transform axis inside the sub-problem
    int log_tw;        // tile width
    uint32_t n_valid;  // entries along the transform axis that exist in src; the rest read as 0
    int inverse;
    uint32_t scale;    // Montgomery factor applied on store (bb::R1 = none)
    int log_r1, log_r2;  // final pass: bits of the two leading output digits (0,0 = single pass)
    int chunk_l;       // >= 0: src holds evaluations; apply the 2^chunk_l chunk iDFT + zeta on load
    const uint32_t* tw_lo;  // (scaled by 2^-chunk_l when chunk_l >= 0)
    const uint32_t* tw_lo_plain;
    const uint32_t* tw_hi;
};

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return bits ? __brev(x) >> (32 - bits) : 0u; }

// 2^G-point decimation-in-frequency butterfly network on registers.  The 2^G values are the
// entries r0 + i*2^b0 (i < 2^G) of a length-2^log_r transform whose bits above b0+G have already
// been processed; `lo` = r0 mod 2^b0.  Natural order in, local bit-reversed order out (in place).
template <int G>
__device__ __forceinline__ void dif_regs(uint32_t (&x)[1 << G], const uint32_t* __restrict__ tw_hi, uint32_t lo,
                                         int b0, int inverse) {
#pragma unroll
    for (int t = 0; t < G; t++) {
        constexpr int dummy = 0;
        (void)dummy;
        const int half = 1 << (G - 1 - t);
        const int shift = TW_HI_BITS - (b0 + G - t);  // twiddle = w_L^(a*2^b0 + lo), L = 2^(b0+G-t)
#pragma unroll
        for (int a = 0; a < half; a++) {
            uint32_t e = (((uint32_t)a << b0) + lo) << shift;
            if (inverse) e = ((1u << TW_HI_BITS) - e) & TW_HI_MASK;
            // lowest group of a transform (b0 == 0, so lo == 0): the a == 0 butterflies have twiddle w^0 = 1 -- 7 of
            // the 12 products of a radix-8 group, 15 of 32 of a radix-16 group -- and need no multiplication
            const bool unit = b0 == 0 && a == 0;
            const uint32_t w = unit ? bb::R1 : __ldg(tw_hi + e);
#pragma unroll
            for (int blk = 0; blk < (1 << t); blk++) {
                const int i0 = blk * 2 * half + a, i1 = i0 + half;
                const uint32_t u = x[i0], v = x[i1];
                x[i0] = bb::add(u, v);
                x[i1] = unit ? bb::sub(u, v) : bb::mul_diff(u, v, w);
            }
        }
    }
}

// One butterfly group over the shared tile sm[r * pitch + s].  FINAL: hand the 2^G results of a
// task to `out(r0, s, x, integral_constant<G>)` (position of x[i] is r0 + i, b0 == 0) instead of
// writing them back.  All arguments are compile-time constants in the specialised kernels.
template <int G, bool FINAL, class Out>
__device__ __forceinline__ void tile_group(uint32_t* sm, int log_r, int b0, int log_tw, int pitch, int inverse,
                                           const uint32_t* __restrict__ tw_hi, Out out) {
    const int tasks = (1 << (log_r - G)) << log_tw;
#pragma unroll
    for (int idx = threadIdx.x; idx < tasks; idx += NTT_THREADS) {
        const uint32_t s = idx & ((1u << log_tw) - 1), q = (uint32_t)idx >> log_tw;
        const uint32_t lo = q & ((1u << b0) - 1);
        const uint32_t r0 = ((q >> b0) << (b0 + G)) + lo;
        uint32_t* p = sm + r0 * pitch + s;
        uint32_t x[1 << G];
#pragma unroll
        for (int i = 0; i < (1 << G); i++) x[i] = p[(i << b0) * pitch];
        dif_regs<G>(x, tw_hi, lo, b0, inverse);
        if (FINAL) {
            out(r0, s, x, std::integral_constant<int, G>{});
        } else {
#pragma unroll
            for (int i = 0; i < (1 << G); i++) p[(i << b0) * pitch] = x[i];
        }
    }
}

template <bool FINAL, class Out>
__device__ __forceinline__ void tile_group_dyn(int G, uint32_t* sm, int log_r, int b0, int log_tw, int pitch,
                                               int inverse, const uint32_t* __restrict__ tw_hi, Out out) {
    switch (G) {
        case 1: tile_group<1, FINAL>(sm, log_r, b0, log_tw, pitch, inverse, tw_hi, out); break;
        case 2: tile_group<2, FINAL>(sm, log_r, b0, log_tw, pitch, inverse, tw_hi, out); break;
        case 3: tile_group<3, FINAL>(sm, log_r, b0, log_tw, pitch, inverse, tw_hi, out); break;
        default: tile_group<4, FINAL>(sm, log_r, b0, log_tw, pitch, inverse, tw_hi, out); break;
    }
}

// Compile-time schedule: HI_BIT index bits left, in GROUPS_LEFT balanced groups (10 -> 4,3,3).
template <int LR, int LT, int HI_BIT, int GROUPS_LEFT, class Out>
__device__ __forceinline__ void tile_dif_static(uint32_t* sm, int inverse, const uint32_t* __restrict__ tw_hi, Out out) {
    constexpr int G = (HI_BIT + GROUPS_LEFT - 1) / GROUPS_LEFT;
    constexpr int B0 = HI_BIT - G;
    if constexpr (GROUPS_LEFT == 1) {
        tile_group<G, true>(sm, LR, B0, LT, (1 << LT) + 1, inverse, tw_hi, out);
    } else {
        tile_group<G, false>(sm, LR, B0, LT, (1 << LT) + 1, inverse, tw_hi, out);
        __syncthreads();
        tile_dif_static<LR, LT, B0, GROUPS_LEFT - 1>(sm, inverse, tw_hi, out);
    }
}

// Length-2^log_r DIF transform of every lane of the tile: groups of <= 4 index bits are done in
// registers, with one shared-memory exchange between groups.  Natural order in; the value at
// position r that reaches `out` is output number bitrev(r).  LR > 0: everything static.
template <int LR, int LT, class Out>
__device__ __forceinline__ void tile_dif(uint32_t* sm, int log_r, int log_tw, int pitch, int inverse,
                                         const uint32_t* __restrict__ tw_hi, Out out) {
    if constexpr (LR > 0) {
        tile_dif_static<LR, LT, LR, (LR + 3) / 4>(sm, inverse, tw_hi, out);
    } else {
        const int ngroups = (log_r + 3) >> 2;
        if (ngroups == 0) {  // length-1 transform
            for (int s = threadIdx.x; s < (1 << log_tw); s += NTT_THREADS) {
                uint32_t x[1] = {sm[s]};
                out(0u, (uint32_t)s, x, std::integral_constant<int, 0>{});
            }
            return;
        }
        int hi_bit = log_r;
        for (int g = 0; g < ngroups; g++) {
            const int left = ngroups - g;
            const int G = (hi_bit + left - 1) / left;
            const int b0 = hi_bit - G;
            if (g == ngroups - 1) {
                tile_group_dyn<true>(G, sm, log_r, b0, log_tw, pitch, inverse, tw_hi, out);
            } else {
                tile_group_dyn<false>(G, sm, log_r, b0, log_tw, pitch, inverse, tw_hi, out);
                __syncthreads();
            }
            hi_bit = b0;
        }
    }
}

template <int G>
__device__ __forceinline__ constexpr int bitrev_c(int i) {
    int r = 0;
    for (int b = 0; b < G; b++) r |= ((i >> b) & 1) << (G - 1 - b);
    return r;
}

// In registers: for every 2^L chunk of the 16 values, inverse DFT (without the 1/2^L factor, which
// is folded into the inter-pass twiddle) followed by the subset-zeta transform (poly.rs:325-348).
template <int L>
__device__ __forceinline__ void chunk16(uint32_t (&x)[16], const uint32_t* __restrict__ tw_hi) {
    if (L == 0) return;
#pragma unroll
    for (int t = 0; t < L; t++) {
        const int half = 1 << (L - 1 - t);
#pragma unroll
        for (int a = 0; a < half; a++) {
            const uint32_t e = ((1u << TW_HI_BITS) - ((uint32_t)a << (TW_HI_BITS - (L - t)))) & TW_HI_MASK;
            const uint32_t w = __ldg(tw_hi + e);
#pragma unroll
            for (int blk = 0; blk < (16 >> (L - t)); blk++) {
                const int i0 = blk * 2 * half + a, i1 = i0 + half;
                const uint32_t u = x[i0], v = x[i1];
                x[i0] = bb::add(u, v);
                x[i1] = bb::mul_diff(u, v, w);
            }
        }
    }
    // zeta commutes across index bits, so it runs on the bit-reversed layout left by the DIF
#pragma unroll
    for (int bit = 0; bit < L; bit++)
#pragma unroll
        for (int u = 0; u < 16; u++)
            if (!(u & (1 << bit))) x[u + (1 << bit)] = bb::add(x[u + (1 << bit)], x[u]);
    uint32_t y[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int c = i >> L, k = i & ((1 << L) - 1);
        int rk = 0;
#pragma unroll
        for (int b = 0; b < L; b++) rk |= ((k >> b) & 1) << (L - 1 - b);
        y[i] = x[(c << L) + rk];
    }
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = y[i];
}

// Non-final pass: data viewed as [outer][R][S]; tile = all R rows x TW consecutive s.
// LR > 0: radix and tile width fixed at compile time (host guarantees a.log_r == LR, a.log_tw == LT).
template <int LR, int LT>
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_strided_pass_kernel(PassArgs a) {
    extern __shared__ uint32_t sm[];
    const int log_r = LR > 0 ? LR : a.log_r, log_tw = LR > 0 ? LT : a.log_tw;
    const int R = 1 << log_r, TW = 1 << log_tw;
    const int pitch = TW + 1;
    const int log_m = log_r + a.log_s;
    const size_t tiles_per_col = size_t(1) << (a.log_n - log_r - log_tw);
    const size_t col = blockIdx.x / tiles_per_col, tile = blockIdx.x % tiles_per_col;
    const size_t o = tile >> (a.log_s - log_tw);
    const uint32_t s0 = (uint32_t)(tile & ((size_t(1) << (a.log_s - log_tw)) - 1)) << log_tw;
    const size_t base = (o << log_m) + s0;
    const uint32_t* src = a.src + col * a.src_col_stride + base;
    uint32_t* dst = a.dst + col * a.dst_col_stride + base;

    if (a.chunk_l >= 0) {
        // TW >= 16: a thread owns 16 consecutive evaluations of a row = whole chunks
        const int log_parts = log_tw - 4;
        for (int idx = threadIdx.x; idx < (R << log_parts); idx += NTT_THREADS) {
            const int r = idx >> log_parts, part = idx & ((1 << log_parts) - 1);
            uint32_t x[16];
            if ((uint32_t)r < a.n_valid) {
                const uint4* p = reinterpret_cast<const uint4*>(src + ((size_t)r << a.log_s) + (part << 4));
                const uint4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
                x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
                x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                x[8] = v2.x; x[9] = v2.y; x[10] = v2.z; x[11] = v2.w;
                x[12] = v3.x; x[13] = v3.y; x[14] = v3.z; x[15] = v3.w;
                switch (a.chunk_l) {
                    case 1: chunk16<1>(x, a.tw_hi); break;
                    case 2: chunk16<2>(x, a.tw_hi); break;
                    case 3: chunk16<3>(x, a.tw_hi); break;
                    case 4: chunk16<4>(x, a.tw_hi); break;
                    default: break;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) x[i] = 0;
            }
#pragma unroll
            for (int i = 0; i < 16; i++) sm[r * pitch + (part << 4) + i] = x[i];
        }
    } else {
        for (int idx = threadIdx.x; idx < (R << log_tw); idx += NTT_THREADS) {
            const uint32_t r = idx >> log_tw, s = idx & (TW - 1);
            sm[r * pitch + s] = r < a.n_valid ? __ldg(src + ((size_t)r << a.log_s) + s) : 0u;
        }
    }
    __syncthreads();
    const bool scaled = a.chunk_l >= 0;
    tile_dif<LR, LT>(sm, log_r, log_tw, pitch, a.inverse, a.tw_hi, [&](uint32_t r0, uint32_t s, uint32_t* x, auto gt) {
        constexpr int G = decltype(gt)::value;
        // outputs k_i = kq + bitrev_G(i) * 2^(log_r-G); inter-pass twiddle w_M^((s0+s) * k_i)
        //   = w0 * beta^bitrev_G(i),  w0 = w_M^((s0+s) kq),  beta = w_{2^(log_s+G)}^(s0+s)
        const uint32_t kq = bitrev(r0, log_r);
        const uint32_t c = s0 + s;
        uint32_t E0 = (uint32_t)(((uint64_t)c * kq) << (27 - log_m)) & W27_MASK;
        uint32_t D = (uint32_t)((uint64_t)c << (27 - a.log_s - G)) & W27_MASK;
        if (a.inverse) {
            E0 = ((1u << 27) - E0) & W27_MASK;
            D = ((1u << 27) - D) & W27_MASK;
        }
        uint32_t w[1 << G];
        w[0] = scaled ? root_pow<true>(a.tw_lo, a.tw_hi, E0) : root_pow<false>(a.tw_lo, a.tw_hi, E0);
        if (G > 0) {
            uint32_t bp = root_pow<false>(scaled ? a.tw_lo_plain : a.tw_lo, a.tw_hi, D);
#pragma unroll
            for (int lvl = 0; lvl < G; lvl++) {
#pragma unroll
                for (int j = 0; j < (1 << lvl); j++) w[(1 << lvl) + j] = bb::mul(w[j], bp);
                if (lvl + 1 < G) bp = bb::sqr(bp);
            }
        }
        uint32_t* d = dst + ((size_t)kq << a.log_s) + s;
#pragma unroll
        for (int i = 0; i < (1 << G); i++) {
            constexpr int dummy = 0;
            (void)dummy;
            const int j = bitrev_c<G>(i);
            d[(size_t)j << (log_r - G + a.log_s)] = bb::mul(x[i], w[j]);
        }
    });
}

// Final pass: rows of R contiguous elements; TW rows per tile, stored at natural positions.
template <int LR, int LT>
__global__ void __launch_bounds__(NTT_THREADS, 2) ntt_final_pass_kernel(PassArgs a) {
    extern __shared__ uint32_t sm[];
    const int log_r = LR > 0 ? LR : a.log_r, log_tw = LR > 0 ? LT : a.log_tw;
    const int R = 1 << log_r, TW = 1 << log_tw;
    const int pitch = TW + 1;
    const bool single = (a.log_r1 == 0 && a.log_r2 == 0);
    size_t src_row0, src_row_step, dst0, dst_row_step, dst_k_step;
    uint32_t rows_live = TW;
    if (single) {
        // one row per column: the tile spans TW consecutive columns
        const size_t col0 = (size_t)blockIdx.x << log_tw;
        rows_live = (uint32_t)min((size_t)TW, (size_t)a.cols - col0);
        src_row0 = col0 * a.src_col_stride;
        src_row_step = a.src_col_stride;
        dst0 = col0 * a.dst_col_stride;
        dst_row_step = a.dst_col_stride;
        dst_k_step = 1;
    } else {
        const size_t tiles_per_col = size_t(1) << (a.log_r1 + a.log_r2 - log_tw);
        const size_t col = blockIdx.x / tiles_per_col, tile = blockIdx.x % tiles_per_col;
        const size_t k2 = tile & ((size_t(1) << a.log_r2) - 1);
        const size_t k1_0 = (tile >> a.log_r2) << log_tw;
        // in-place row index o = k1 * R2 + k2 ; natural prefix q = k1 + R1 * k2
        src_row0 = col * a.src_col_stride + (((k1_0 << a.log_r2) + k2) << log_r);
        src_row_step = size_t(1) << (a.log_r2 + log_r);
        dst0 = col * a.dst_col_stride + k1_0 + (k2 << a.log_r1);
        dst_row_step = 1;
        dst_k_step = size_t(1) << (a.log_r1 + a.log_r2);
    }
#pragma unroll 8
    for (int idx = threadIdx.x; idx < (TW << log_r); idx += NTT_THREADS) {
        const uint32_t row = idx >> log_r, j = idx & (R - 1);
        uint32_t v = 0;
        if (row < rows_live && j < a.n_valid) v = __ldg(a.src + src_row0 + row * src_row_step + j);
        sm[j * pitch + row] = v;
    }
    __syncthreads();
    const bool scaled = a.scale != bb::R1;
    tile_dif<LR, LT>(sm, log_r, log_tw, pitch, a.inverse, a.tw_hi, [&](uint32_t r0, uint32_t row, uint32_t* x, auto gt) {
        constexpr int G = decltype(gt)::value;
        if (row >= rows_live) return;
        const uint32_t kq = bitrev(r0, log_r);
        uint32_t* d = a.dst + dst0 + row * dst_row_step + kq * dst_k_step;
#pragma unroll
        for (int i = 0; i < (1 << G); i++) {
            const int j = bitrev_c<G>(i);
            uint32_t v = x[i];
            if (scaled) v = bb::mul(v, a.scale);
            d[((size_t)j << (log_r - G)) * dst_k_step] = v;
        }
    });
}

// Per 2^l chunk: inverse DFT, then subset-zeta over the l index bits (poly.rs:325-348).
// A CTA handles `1 << log_te` consecutive elements (whole chunks) of one column.
__global__ void __launch_bounds__(256)
chunk_coeffs_kernel(const uint32_t* __restrict__ src, size_t src_col_stride, uint32_t* __restrict__ dst,
                    size_t dst_col_stride, int log_h, int l, int log_te, uint32_t scale,
                    const uint32_t* __restrict__ tw_hi) {
    extern __shared__ uint32_t sm[];
    const int TE = 1 << log_te;
    const size_t tiles_per_col = size_t(1) << (log_h - log_te);
    const size_t col = blockIdx.x / tiles_per_col, tile = blockIdx.x % tiles_per_col;
    const uint32_t* s = src + col * src_col_stride + (tile << log_te);
    uint32_t* d = dst + col * dst_col_stride + (tile << log_te);
    for (int i = threadIdx.x; i < TE; i += blockDim.x) sm[i] = __ldg(s + i);
    __syncthreads();
    // DIF with inverse roots inside every chunk (chunk c occupies sm[c<<l .. )
    for (int t = 0; t < l; t++) {
        const int log_half = l - 1 - t;
        for (int idx = threadIdx.x; idx < (TE >> 1); idx += blockDim.x) {
            const int c = idx >> (l - 1), b = idx & ((1 << (l - 1)) - 1);
            const int pos = b & ((1 << log_half) - 1);
            const int i0 = (c << l) + ((b >> log_half) << (log_half + 1)) + pos;
            const int i1 = i0 + (1 << log_half);
            uint32_t e = (uint32_t)pos << (t + TW_HI_BITS - l);
            e = ((1u << TW_HI_BITS) - e) & ((1u << TW_HI_BITS) - 1);
            const uint32_t w = __ldg(tw_hi + e);
            const uint32_t u = sm[i0], v = sm[i1];
            sm[i0] = bb::add(u, v);
            sm[i1] = bb::mul(bb::sub(u, v), w);
        }
        __syncthreads();
    }
    // zeta: a[v] += a[u] for every index bit; the bits commute, so the bit-reversed layout left
    // by the DIF is handled by simply walking position bits.
    for (int bit = 0; bit < l; bit++) {
        for (int idx = threadIdx.x; idx < (TE >> 1); idx += blockDim.x) {
            const int lo = idx & ((1 << bit) - 1);
            const int u = ((idx >> bit) << (bit + 1)) + lo;  // bit `bit` clear (stays inside the chunk)
            sm[u + (1 << bit)] = bb::add(sm[u + (1 << bit)], sm[u]);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < TE; i += blockDim.x) {
        const int c = i >> l, k = i & ((1 << l) - 1);
        d[i] = bb::mul(sm[(c << l) + (int)bitrev(k, l)], scale);
    }
}

// ------------------------------------------------------------------------------------------
// host: planning + launch
// ------------------------------------------------------------------------------------------
struct NttPlan {
    int m = 0;
    int r[3] = {0, 0, 0};
};

static int make_plan(const swirl_ctx* ctx, int log_n, int min_r1, NttPlan* plan) {
    const int max_r = ctx->ntt_max_log_radix;
    int m = log_n <= max_r ? 1 : (log_n + max_r - 1) / max_r;
    if (m > 3) {
        set_error("NTT size needs more than three passes");
        return SWIRL_ERR_INVALID;
    }
    plan->m = m;
    int left = log_n;
    for (int i = 0; i < m; i++) {
        int r = left / (m - i);  // smaller radices first: the first pass then affords 16-wide tiles
        if (i == 0 && r < min_r1) r = std::min(min_r1, log_n);
        plan->r[i] = r;
        left -= r;
    }
    if (plan->r[0] > 13 || left != 0) {
        set_error("unsupported NTT plan");
        return SWIRL_ERR_INVALID;
    }
    // a first-pass radix forced up may leave nothing for later passes
    int mm = 0;
    for (int i = 0; i < m; i++)
        if (plan->r[i] > 0) plan->r[mm++] = plan->r[i];
    for (int i = mm; i < 3; i++) plan->r[i] = 0;
    plan->m = mm ? mm : 1;
    return 0;
}

static int tile_log_tw(int log_r, int log_limit) {
    int log_tw = ilog2(NTT_TILE_ELEMS) - log_r;
    // small radices (the passes of a three-pass plan, 2^24 and up) take wider tiles: the same 16 K elements per CTA as a
    // radix-2^10 pass, 128-256 B segments per row
    if (log_tw > 6) log_tw = 6;
    if (log_tw > log_limit) log_tw = log_limit;
    if (log_tw < 0) log_tw = 0;
    return log_tw;
}

static int ensure_smem_attr() {
    static bool done = false;
    if (!done) {
        const int bytes = (NTT_TILE_ELEMS + (NTT_TILE_ELEMS >> 0)) * 4;  // tile + padding, up to R = 2^13
#define SWIRL_NTT_ATTR(LR, LT)                                                                                    \
    SWIRL_CUDA(cudaFuncSetAttribute(ntt_strided_pass_kernel<LR, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
    SWIRL_CUDA(cudaFuncSetAttribute(ntt_final_pass_kernel<LR, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        SWIRL_NTT_ATTR(0, 0)
        SWIRL_NTT_ATTR(6, 6)
        SWIRL_NTT_ATTR(7, 6)
        SWIRL_NTT_ATTR(8, 6)
        SWIRL_NTT_ATTR(9, 5)
        SWIRL_NTT_ATTR(7, 4)
        SWIRL_NTT_ATTR(8, 4)
        SWIRL_NTT_ATTR(9, 4)
        SWIRL_NTT_ATTR(10, 4)
        SWIRL_NTT_ATTR(11, 3)
#undef SWIRL_NTT_ATTR
        done = true;
    }
    return 0;
}

// Transform `cols` columns: src (column stride src_stride, `n_valid` leading entries per column,
// the rest implicitly zero) -> dst (column stride dst_stride) of length 2^log_n, natural order.
// `tmp` (>= cols << log_n words) is needed when the plan has more than one pass.
// chunk_l >= 0 asks the first (strided) pass to apply the 2^chunk_l chunk iDFT + zeta while loading.
static bool can_fuse_chunks(const NttPlan& plan, int log_n, int l_skip, const uint32_t* src, size_t src_stride) {
    if (plan.m < 2 || l_skip < 1 || l_skip > 4) return false;
    const int log_s = log_n - plan.r[0];
    return log_s >= 4 && plan.r[0] <= ilog2(NTT_TILE_ELEMS) - 4 && (src_stride & 3) == 0 &&
           ((uintptr_t)src & 15) == 0;
}

static int run_ntt(swirl_ctx* ctx, const NttPlan& plan, const uint32_t* src, size_t src_stride, uint32_t* dst,
                   size_t dst_stride, uint32_t* tmp, size_t cols, int log_n, size_t n_valid, bool inverse,
                   uint32_t scale, int chunk_l = -1) {
    SWIRL_TRY(ensure_smem_attr());
    PassArgs a{};
    a.cols = (uint32_t)cols;
    a.log_n = log_n;
    a.inverse = inverse ? 1 : 0;
    a.tw_lo = ctx->tw_lo;
    a.tw_hi = ctx->tw_hi;
    const size_t N = size_t(1) << log_n;
    int consumed = 0;
    const uint32_t* cur_src = src;
    size_t cur_stride = src_stride;
    for (int i = 0; i < plan.m - 1; i++) {
        a.log_r = plan.r[i];
        a.log_s = log_n - consumed - a.log_r;
        a.log_tw = tile_log_tw(a.log_r, a.log_s);
        a.src = cur_src;
        a.src_col_stride = cur_stride;
        a.dst = tmp;
        a.dst_col_stride = N;
        a.n_valid = (i == 0) ? (uint32_t)(n_valid >> a.log_s) : (1u << a.log_r);
        a.scale = bb::R1;
        a.chunk_l = (i == 0) ? chunk_l : -1;
        a.tw_lo = ctx->tw_lo;
        if (a.chunk_l >= 0) SWIRL_TRY(scaled_twiddles(ctx, chunk_l, &a.tw_lo));
        const size_t grid = cols << (log_n - a.log_r - a.log_tw);
        SWIRL_REQUIRE(grid < (size_t(1) << 31), "NTT grid too large");
        const size_t smem = (size_t(4) << a.log_r) * ((size_t(1) << a.log_tw) + 1);
        {
            SwirlTimed timed(ctx, SWIRL_T_NTT_PASS);
            a.tw_lo_plain = ctx->tw_lo;
#define SWIRL_NTT_CASE(LR, LT)                                                                  \
    if (a.log_r == LR && a.log_tw == LT)                                                        \
        ntt_strided_pass_kernel<LR, LT><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(a); \
    else
            SWIRL_NTT_CASE(6, 6) SWIRL_NTT_CASE(7, 6) SWIRL_NTT_CASE(8, 6) SWIRL_NTT_CASE(9, 5)
            SWIRL_NTT_CASE(7, 4) SWIRL_NTT_CASE(8, 4) SWIRL_NTT_CASE(9, 4) SWIRL_NTT_CASE(10, 4) SWIRL_NTT_CASE(11, 3)
                ntt_strided_pass_kernel<0, 0><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(a);
#undef SWIRL_NTT_CASE
        }
        SWIRL_LAUNCH_CHECK(ctx);
        consumed += a.log_r;
        cur_src = tmp;
        cur_stride = N;
    }
    // final pass
    a.chunk_l = -1;
    a.tw_lo = ctx->tw_lo;
    a.log_r = plan.r[plan.m - 1];
    a.log_s = 0;
    a.src = cur_src;
    a.src_col_stride = cur_stride;
    a.dst = dst;
    a.dst_col_stride = dst_stride;
    a.scale = scale;
    a.log_r1 = plan.m >= 2 ? plan.r[0] : 0;
    a.log_r2 = plan.m == 3 ? plan.r[1] : 0;
    a.n_valid = plan.m == 1 ? (uint32_t)n_valid : (1u << a.log_r);
    size_t grid;
    if (plan.m == 1) {
        a.log_tw = tile_log_tw(a.log_r, 5);
        grid = (cols + (size_t(1) << a.log_tw) - 1) >> a.log_tw;
    } else {
        a.log_tw = tile_log_tw(a.log_r, a.log_r1);
        grid = cols << (a.log_r1 + a.log_r2 - a.log_tw);
    }
    SWIRL_REQUIRE(grid < (size_t(1) << 31), "NTT grid too large");
    const size_t smem = (size_t(4) << a.log_r) * ((size_t(1) << a.log_tw) + 1);
    {
        SwirlTimed timed(ctx, SWIRL_T_NTT_FINAL);
#define SWIRL_NTT_CASE(LR, LT)                                                                \
    if (a.log_r == LR && a.log_tw == LT)                                                      \
        ntt_final_pass_kernel<LR, LT><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(a); \
    else
        SWIRL_NTT_CASE(6, 6) SWIRL_NTT_CASE(7, 6) SWIRL_NTT_CASE(8, 6) SWIRL_NTT_CASE(9, 5)
        SWIRL_NTT_CASE(7, 4) SWIRL_NTT_CASE(8, 4) SWIRL_NTT_CASE(9, 4) SWIRL_NTT_CASE(10, 4) SWIRL_NTT_CASE(11, 3)
            ntt_final_pass_kernel<0, 0><<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(a);
#undef SWIRL_NTT_CASE
    }
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}

static size_t group_cols(const swirl_ctx* ctx, size_t cols, int log_n, int passes) {
    if (passes <= 1) return cols;
    size_t per_col = size_t(4) << log_n;
    size_t g = ctx->ntt_scratch_bytes / per_col;
    if (g < 1) g = 1;
    return g < cols ? g : cols;
}

int ntt_batch(swirl_ctx* ctx, uint32_t* d_data, int log_n, size_t cols, bool inverse) {
    SWIRL_REQUIRE(log_n >= 0 && log_n <= 27, "log_n must be in [0, 27]");
    if (cols == 0 || log_n == 0) return 0;
    NttPlan plan;
    SWIRL_TRY(make_plan(ctx, log_n, 0, &plan));
    const size_t N = size_t(1) << log_n;
    const uint32_t scale = inverse ? bb::inv(bb::to_mont((uint32_t)(N % bb::P))) : bb::R1;
    const size_t g = group_cols(ctx, cols, log_n, plan.m);
    uint32_t* tmp = nullptr;
    if (plan.m > 1) SWIRL_CUDA(dev_alloc(ctx, &tmp, g << log_n));
    int rc = 0;
    for (size_t c0 = 0; c0 < cols && rc == 0; c0 += g) {
        const size_t nc = std::min(g, cols - c0);
        uint32_t* p = d_data + c0 * N;
        rc = run_ntt(ctx, plan, p, N, p, N, tmp, nc, log_n, N, inverse, scale);
    }
    dev_free(ctx, tmp);
    return rc;
}

int chunk_coeffs(swirl_ctx* ctx, const uint32_t* src, size_t src_stride, uint32_t* dst, size_t dst_stride, size_t H,
                 size_t cols, int l_skip) {
    SWIRL_REQUIRE(is_pow2(H), "height must be a power of two");
    const int log_h = ilog2(H);
    SWIRL_REQUIRE(l_skip >= 0 && l_skip <= log_h && l_skip <= 11, "l_skip");
    if (cols == 0) return 0;
    int log_te = std::min(log_h, 11);
    if (log_te < l_skip) log_te = l_skip;
    const size_t grid = cols << (log_h - log_te);
    SWIRL_REQUIRE(grid < (size_t(1) << 31), "grid too large");
    chunk_coeffs_kernel<<<(unsigned)grid, 256, size_t(4) << log_te, ctx->stream>>>(
        src, src_stride, dst, dst_stride, log_h, l_skip, log_te, bb::inv(bb::to_mont(1u << l_skip)), ctx->tw_hi);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}

int rs_encode(swirl_ctx* ctx, const uint32_t* d_in, size_t in_stride, size_t H, size_t W, int l_skip,
              int log_blowup, uint32_t* d_out) {
    SWIRL_REQUIRE(is_pow2(H), "stacked height must be a power of two");
    const int log_h = ilog2(H);
    SWIRL_REQUIRE(l_skip >= 0 && l_skip <= log_h, "l_skip exceeds log height");
    SWIRL_REQUIRE(l_skip <= 11, "l_skip > 11 unsupported");
    const int log_n = log_h + log_blowup;
    SWIRL_REQUIRE(log_blowup >= 0 && log_n <= 27, "codeword longer than 2^27");
    if (W == 0) return 0;
    const size_t N = size_t(1) << log_n;
    NttPlan plan;
    SWIRL_TRY(make_plan(ctx, log_n, log_blowup, &plan));
    const bool fuse = can_fuse_chunks(plan, log_n, l_skip, d_in, in_stride);
    // group columns so that message scratch + pass scratch stay cache resident
    size_t g = W;
    {
        size_t per_col = (plan.m > 1 ? (size_t(4) << log_n) : 0) + (l_skip > 0 && !fuse ? (size_t(4) << log_h) : 0);
        if (per_col) {
            g = ctx->ntt_scratch_bytes / per_col;
            if (g < 1) g = 1;
            if (g > W) g = W;
        }
    }
    uint32_t *msg = nullptr, *tmp = nullptr;
    if (l_skip > 0 && !fuse) SWIRL_CUDA(dev_alloc(ctx, &msg, g << log_h));
    if (plan.m > 1) SWIRL_CUDA(dev_alloc(ctx, &tmp, g << log_n));
    const uint32_t chunk_scale = bb::inv(bb::to_mont(1u << l_skip));
    int rc = 0;
    for (size_t c0 = 0; c0 < W && rc == 0; c0 += g) {
        const size_t nc = std::min(g, W - c0);
        const uint32_t* src = d_in + c0 * in_stride;
        size_t src_stride = in_stride;
        if (l_skip > 0 && !fuse) {
            int log_te = std::min(log_h, 11);
            if (log_te < l_skip) log_te = l_skip;
            const size_t grid = nc << (log_h - log_te);
            {
                SwirlTimed timed(ctx, SWIRL_T_CHUNK);
                chunk_coeffs_kernel<<<(unsigned)grid, 256, size_t(4) << log_te, ctx->stream>>>(
                    src, in_stride, msg, H, log_h, l_skip, log_te, chunk_scale, ctx->tw_hi);
            }
            ctx->launches++;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) {
                rc = cuda_fail(e, "chunk_coeffs_kernel", __FILE__, __LINE__);
                break;
            }
            src = msg;
            src_stride = H;
        }
        rc = run_ntt(ctx, plan, src, src_stride, d_out + c0 * N, N, tmp, nc, log_n, H, false, bb::R1,
                     fuse ? l_skip : -1);
    }
    dev_free(ctx, msg);
    dev_free(ctx, tmp);
    return rc;
}

}  // namespace swirl
