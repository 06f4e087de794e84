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
#include <vector>

#include "bb31.cuh"
#include "kernels.cuh"

namespace swirl {

constexpr int NTT_THREADS = 256;
constexpr int NTT_TILE_ELEMS = 16384;  // words of shared memory per tile (before padding)
constexpr uint32_t W27_MASK = (1u << 27) - 1;

// w^E for the 2^27-th root w, E < 2^27, via two tables (one multiply unless E's low bits vanish)
__device__ __forceinline__ uint32_t root_pow(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi,
                                             uint32_t E) {
    uint32_t h = __ldg(hi + (E >> TW_LO_BITS));
    uint32_t l = E & ((1u << TW_LO_BITS) - 1);
    return l ? bb::mul(h, __ldg(lo + l)) : h;
}

__global__ void twiddle_init_kernel(uint32_t* lo, uint32_t* hi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = bb::two_adic_generator(27);
    if (i < (1u << TW_LO_BITS)) lo[i] = bb::pow(w, i);
    if (i < (1u << TW_HI_BITS)) hi[i] = bb::pow(w, (uint64_t)i << TW_LO_BITS);
}

int ntt_init_twiddles(swirl_ctx* ctx) {
    SWIRL_CUDA(cudaMalloc((void**)&ctx->tw_lo, sizeof(uint32_t) << TW_LO_BITS));
    SWIRL_CUDA(cudaMalloc((void**)&ctx->tw_hi, sizeof(uint32_t) << TW_HI_BITS));
    twiddle_init_kernel<<<(1u << TW_LO_BITS) / 256, 256, 0, ctx->stream>>>(ctx->tw_lo, ctx->tw_hi);
    SWIRL_LAUNCH_CHECK(ctx);
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
    const uint32_t* tw_lo;
    const uint32_t* tw_hi;
};

// In-place radix-2 DIF along the slow axis of a [R][pitch] shared tile; natural order in,
// bit-reversed order out.  All `width` lanes of a row are transformed independently.
__device__ __forceinline__ void tile_dif(uint32_t* sm, int log_r, int log_w, int pitch, int inverse,
                                         const uint32_t* __restrict__ tw_hi) {
    const int R = 1 << log_r;
    const int total = (R >> 1) << log_w;
    for (int t = 0; t < log_r; t++) {
        const int log_half = log_r - 1 - t;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int b = idx >> log_w, s = idx & ((1 << log_w) - 1);
            const int pos = b & ((1 << log_half) - 1);
            const int i0 = ((b >> log_half) << (log_half + 1)) + pos;
            const int i1 = i0 + (1 << log_half);
            // w_R^(pos << t) = w_{2^13}^(pos << (t + 13 - log_r))
            uint32_t e = (uint32_t)pos << (t + TW_HI_BITS - log_r);
            if (inverse) e = ((1u << TW_HI_BITS) - e) & ((1u << TW_HI_BITS) - 1);
            const uint32_t w = __ldg(tw_hi + e);
            const uint32_t u = sm[i0 * pitch + s], v = sm[i1 * pitch + s];
            sm[i0 * pitch + s] = bb::add(u, v);
            sm[i1 * pitch + s] = bb::mul(bb::sub(u, v), w);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return bits ? __brev(x) >> (32 - bits) : 0u; }

// Non-final pass: data viewed as [outer][R][S]; tile = all R rows x TW consecutive s.
__global__ void __launch_bounds__(NTT_THREADS) ntt_strided_pass_kernel(PassArgs a) {
    extern __shared__ uint32_t sm[];
    const int R = 1 << a.log_r, TW = 1 << a.log_tw;
    const int log_m = a.log_r + a.log_s;
    const size_t tiles_per_col = size_t(1) << (a.log_n - a.log_r - a.log_tw);
    const size_t col = blockIdx.x / tiles_per_col, tile = blockIdx.x % tiles_per_col;
    const size_t o = tile >> (a.log_s - a.log_tw);
    const uint32_t s0 = (uint32_t)(tile & ((size_t(1) << (a.log_s - a.log_tw)) - 1)) << a.log_tw;
    const size_t base = (o << log_m) + s0;
    const uint32_t* src = a.src + col * a.src_col_stride + base;
    uint32_t* dst = a.dst + col * a.dst_col_stride + base;

    for (int idx = threadIdx.x; idx < (R << a.log_tw); idx += blockDim.x) {
        const uint32_t r = idx >> a.log_tw, s = idx & (TW - 1);
        sm[idx] = r < a.n_valid ? __ldg(src + ((size_t)r << a.log_s) + s) : 0u;
    }
    __syncthreads();
    tile_dif(sm, a.log_r, a.log_tw, TW, a.inverse, a.tw_hi);
    for (int idx = threadIdx.x; idx < (R << a.log_tw); idx += blockDim.x) {
        const uint32_t k = idx >> a.log_tw, s = idx & (TW - 1);
        uint32_t v = sm[(bitrev(k, a.log_r) << a.log_tw) + s];
        // inter-pass twiddle w_M^((s0+s) * k)
        uint32_t E = (uint32_t)(((uint64_t)(s0 + s) * k) << (27 - log_m)) & W27_MASK;
        if (a.inverse) E = ((1u << 27) - E) & W27_MASK;
        v = bb::mul(v, root_pow(a.tw_lo, a.tw_hi, E));
        dst[((size_t)k << a.log_s) + s] = v;
    }
}

// Final pass: rows of R contiguous elements; TW rows per tile, stored at natural positions.
__global__ void __launch_bounds__(NTT_THREADS) ntt_final_pass_kernel(PassArgs a) {
    extern __shared__ uint32_t sm[];
    const int R = 1 << a.log_r, TW = 1 << a.log_tw;
    const int pitch = TW + 1;
    const bool single = (a.log_r1 == 0 && a.log_r2 == 0);
    size_t src_row0, src_row_step, dst0, dst_row_step, dst_k_step;
    uint32_t rows_live = TW;
    if (single) {
        // one row per column: the tile spans TW consecutive columns
        const size_t col0 = (size_t)blockIdx.x << a.log_tw;
        rows_live = (uint32_t)min((size_t)TW, (size_t)a.cols - col0);
        src_row0 = col0 * a.src_col_stride;
        src_row_step = a.src_col_stride;
        dst0 = col0 * a.dst_col_stride;
        dst_row_step = a.dst_col_stride;
        dst_k_step = 1;
    } else {
        const size_t tiles_per_col = size_t(1) << (a.log_r1 + a.log_r2 - a.log_tw);
        const size_t col = blockIdx.x / tiles_per_col, tile = blockIdx.x % tiles_per_col;
        const size_t k2 = tile & ((size_t(1) << a.log_r2) - 1);
        const size_t k1_0 = (tile >> a.log_r2) << a.log_tw;
        // in-place row index o = k1 * R2 + k2 ; natural prefix q = k1 + R1 * k2
        src_row0 = col * a.src_col_stride + (((k1_0 << a.log_r2) + k2) << a.log_r);
        src_row_step = size_t(1) << (a.log_r2 + a.log_r);
        dst0 = col * a.dst_col_stride + k1_0 + (k2 << a.log_r1);
        dst_row_step = 1;
        dst_k_step = size_t(1) << (a.log_r1 + a.log_r2);
    }
    for (int idx = threadIdx.x; idx < (TW << a.log_r); idx += blockDim.x) {
        const uint32_t row = idx >> a.log_r, j = idx & (R - 1);
        uint32_t v = 0;
        if (row < rows_live && j < a.n_valid) v = __ldg(a.src + src_row0 + row * src_row_step + j);
        sm[j * pitch + row] = v;
    }
    __syncthreads();
    tile_dif(sm, a.log_r, a.log_tw, pitch, a.inverse, a.tw_hi);
    const bool scaled = a.scale != bb::R1;
    if (single) {
        for (int idx = threadIdx.x; idx < (TW << a.log_r); idx += blockDim.x) {
            const uint32_t row = idx >> a.log_r, k = idx & (R - 1);
            if (row >= rows_live) continue;
            uint32_t v = sm[bitrev(k, a.log_r) * pitch + row];
            if (scaled) v = bb::mul(v, a.scale);
            a.dst[dst0 + row * dst_row_step + k] = v;
        }
    } else {
        for (int idx = threadIdx.x; idx < (TW << a.log_r); idx += blockDim.x) {
            const uint32_t k = idx >> a.log_tw, row = idx & (TW - 1);
            uint32_t v = sm[bitrev(k, a.log_r) * pitch + row];
            if (scaled) v = bb::mul(v, a.scale);
            a.dst[dst0 + row * dst_row_step + k * dst_k_step] = v;
        }
    }
}

// Per 2^l chunk: inverse DFT, then subset-zeta over the l index bits (poly.rs:325-348).
// A CTA handles `1 << log_te` consecutive elements (whole chunks) of one column.
__global__ void __launch_bounds__(NTT_THREADS)
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
        int r = (left + (m - i) - 1) / (m - i);
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
    if (log_tw > 5) log_tw = 5;
    if (log_tw > log_limit) log_tw = log_limit;
    if (log_tw < 0) log_tw = 0;
    return log_tw;
}

static int ensure_smem_attr() {
    static bool done = false;
    if (!done) {
        const int bytes = (NTT_TILE_ELEMS + (NTT_TILE_ELEMS >> 0)) * 4;  // generous: tile + padding
        SWIRL_CUDA(cudaFuncSetAttribute(ntt_strided_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        SWIRL_CUDA(cudaFuncSetAttribute(ntt_final_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done = true;
    }
    return 0;
}

// Transform `cols` columns: src (column stride src_stride, `n_valid` leading entries per column,
// the rest implicitly zero) -> dst (column stride dst_stride) of length 2^log_n, natural order.
// `tmp` (>= cols << log_n words) is needed when the plan has more than one pass.
static int run_ntt(swirl_ctx* ctx, const NttPlan& plan, const uint32_t* src, size_t src_stride, uint32_t* dst,
                   size_t dst_stride, uint32_t* tmp, size_t cols, int log_n, size_t n_valid, bool inverse,
                   uint32_t scale) {
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
        const size_t grid = cols << (log_n - a.log_r - a.log_tw);
        SWIRL_REQUIRE(grid < (size_t(1) << 31), "NTT grid too large");
        const size_t smem = (size_t(4) << a.log_r) << a.log_tw;
        {
            SwirlTimed timed(ctx, SWIRL_T_NTT_PASS);
            ntt_strided_pass_kernel<<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(a);
        }
        SWIRL_LAUNCH_CHECK(ctx);
        consumed += a.log_r;
        cur_src = tmp;
        cur_stride = N;
    }
    // final pass
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
        ntt_final_pass_kernel<<<(unsigned)grid, NTT_THREADS, smem, ctx->stream>>>(a);
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
    // group columns so that message scratch + pass scratch stay cache resident
    size_t g = W;
    {
        size_t per_col = (plan.m > 1 ? (size_t(4) << log_n) : 0) + (l_skip > 0 ? (size_t(4) << log_h) : 0);
        if (per_col) {
            g = ctx->ntt_scratch_bytes / per_col;
            if (g < 1) g = 1;
            if (g > W) g = W;
        }
    }
    uint32_t *msg = nullptr, *tmp = nullptr;
    if (l_skip > 0) SWIRL_CUDA(dev_alloc(ctx, &msg, g << log_h));
    if (plan.m > 1) SWIRL_CUDA(dev_alloc(ctx, &tmp, g << log_n));
    const uint32_t chunk_scale = bb::inv(bb::to_mont(1u << l_skip));
    int rc = 0;
    for (size_t c0 = 0; c0 < W && rc == 0; c0 += g) {
        const size_t nc = std::min(g, W - c0);
        const uint32_t* src = d_in + c0 * in_stride;
        size_t src_stride = in_stride;
        if (l_skip > 0) {
            int log_te = std::min(log_h, 11);
            if (log_te < l_skip) log_te = l_skip;
            const size_t grid = nc << (log_h - log_te);
            {
                SwirlTimed timed(ctx, SWIRL_T_CHUNK);
                chunk_coeffs_kernel<<<(unsigned)grid, NTT_THREADS, size_t(4) << log_te, ctx->stream>>>(
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
        rc = run_ntt(ctx, plan, src, src_stride, d_out + c0 * N, N, tmp, nc, log_n, H, false, bb::R1);
    }
    dev_free(ctx, msg);
    dev_free(ctx, tmp);
    return rc;
}

}  // namespace swirl
