// Multilinear-polynomial building blocks shared by the opening phases: tensor-product tables
// (eq / Möbius-eq kernels on the hypercube) and the subset-sum ("zeta") transform between MLE
// coefficients and hypercube evaluations.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-backend/cuda/src/mle_interpolate.cu:16-445    mle_interpolate_* stages
//   crates/cuda-backend/cuda/src/poly.cu                      eq_hypercube / mobius_eq builders
// Semantics = prover/poly.rs:99-131 (Mle::evals_to_coeffs_inplace / coeffs_to_evals_inplace),
// :133-178 (evals_eq_hypercube, evals_mobius_eq_hypercube).
#include "ext.cuh"
#include "kernels.cuh"

namespace swirl {

using bb::ext_mul;

constexpr int MLE_BLOCK = 256;

// out[i] = prod_b (bit b of i ? w1[b] : w0[b]),  b < n_vars, i < 2^n_vars
__global__ void __launch_bounds__(MLE_BLOCK)
tensor_table_kernel(TensorArgs t, int first_var, int n_vars, uint32_t* __restrict__ out, size_t n_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    Ext acc = bb::ext_one();
    for (int b = 0; b < n_vars; b++) {
        const uint32_t* w = ((i >> b) & 1) ? t.w1[first_var + b] : t.w0[first_var + b];
        acc = ext_mul(acc, Ext{{w[0], w[1], w[2], w[3]}});
    }
    st_ext(out + i * 4, acc);
}

// out[i] = A[i mod 2^lo_bits] * B[i >> lo_bits]
__global__ void __launch_bounds__(MLE_BLOCK)
tensor_combine_kernel(const uint32_t* __restrict__ A, const uint32_t* __restrict__ B, int lo_bits,
                      uint32_t* __restrict__ out, size_t n_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    st_ext(out + i * 4, ext_mul(ldg_ext(A + (i & ((size_t(1) << lo_bits) - 1)) * 4), ldg_ext(B + (i >> lo_bits) * 4)));
}

int mle_tensor_table(swirl_ctx* ctx, const TensorArgs& t, int n_vars, uint32_t* d_out) {
    SWIRL_REQUIRE(n_vars >= 0 && n_vars <= 28, "too many variables");
    const size_t n = size_t(1) << n_vars;
    if (n_vars <= 12) {
        tensor_table_kernel<<<(unsigned)((n + MLE_BLOCK - 1) / MLE_BLOCK), MLE_BLOCK, 0, ctx->stream>>>(t, 0, n_vars, d_out, n);
        SWIRL_LAUNCH_CHECK(ctx);
        return 0;
    }
    const int lo = n_vars / 2, hi = n_vars - lo;
    uint32_t* tmp = nullptr;
    SWIRL_CUDA(dev_alloc(ctx, &tmp, ((size_t(1) << lo) + (size_t(1) << hi)) * 4));
    uint32_t* B = tmp + (size_t(4) << lo);
    tensor_table_kernel<<<(unsigned)(((size_t(1) << lo) + MLE_BLOCK - 1) / MLE_BLOCK), MLE_BLOCK, 0, ctx->stream>>>(
        t, 0, lo, tmp, size_t(1) << lo);
    tensor_table_kernel<<<(unsigned)(((size_t(1) << hi) + MLE_BLOCK - 1) / MLE_BLOCK), MLE_BLOCK, 0, ctx->stream>>>(
        t, lo, hi, B, size_t(1) << hi);
    ctx->launches += 2;
    tensor_combine_kernel<<<(unsigned)((n + MLE_BLOCK - 1) / MLE_BLOCK), MLE_BLOCK, 0, ctx->stream>>>(tmp, B, lo, d_out, n);
    SWIRL_LAUNCH_CHECK(ctx);
    dev_free(ctx, tmp);
    return 0;
}

// ---- zeta transform -----------------------------------------------------------------------------
// Stages for index bits [0, tb): a CTA owns 2^tb consecutive elements of one column in shared memory.
__global__ void __launch_bounds__(MLE_BLOCK)
zeta_tile_kernel(uint32_t* __restrict__ data, size_t col_stride, int log_n, int tb, int inverse) {
    extern __shared__ uint32_t sm[];
    const size_t tiles_per_col = size_t(1) << (log_n - tb);
    const size_t col = blockIdx.x / tiles_per_col, tile = blockIdx.x % tiles_per_col;
    uint32_t* p = data + col * col_stride + (tile << tb);
    const int T = 1 << tb;
    for (int i = threadIdx.x; i < T; i += blockDim.x) sm[i] = p[i];
    __syncthreads();
    for (int bit = 0; bit < tb; bit++) {
        for (int idx = threadIdx.x; idx < (T >> 1); idx += blockDim.x) {
            const int lo = idx & ((1 << bit) - 1);
            const int u = ((idx >> bit) << (bit + 1)) + lo, v = u + (1 << bit);
            sm[v] = inverse ? bb::sub(sm[v], sm[u]) : bb::add(sm[v], sm[u]);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < T; i += blockDim.x) p[i] = sm[i];
}
// One stage for an index bit >= the tile bits: a[v] (+/-)= a[u], 4 consecutive elements per thread.
__global__ void __launch_bounds__(MLE_BLOCK)
zeta_stage_kernel(uint32_t* __restrict__ data, size_t col_stride, int log_n, size_t cols, int bit, int inverse) {
    const size_t per_col = size_t(1) << (log_n - 3);  // quads of pairs
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= per_col * cols) return;
    const size_t col = g / per_col, idx = (g % per_col) << 2;  // idx = pair index, multiple of 4 (bit >= 2)
    const size_t lo = idx & ((size_t(1) << bit) - 1);
    const size_t u = ((idx >> bit) << (bit + 1)) + lo;
    uint4* pu = reinterpret_cast<uint4*>(data + col * col_stride + u);
    uint4* pv = reinterpret_cast<uint4*>(data + col * col_stride + u + (size_t(1) << bit));
    const uint4 a = *pu;
    uint4 b = *pv;
    if (inverse) {
        b.x = bb::sub(b.x, a.x); b.y = bb::sub(b.y, a.y); b.z = bb::sub(b.z, a.z); b.w = bb::sub(b.w, a.w);
    } else {
        b.x = bb::add(b.x, a.x); b.y = bb::add(b.y, a.y); b.z = bb::add(b.z, a.z); b.w = bb::add(b.w, a.w);
    }
    *pv = b;
}

// In place on `cols` columns (column stride col_stride words, 16-byte aligned) of length 2^log_n:
// forward = coeffs_to_evals (a[v] += a[u] for every index bit), inverse = evals_to_coeffs.
int mle_zeta(swirl_ctx* ctx, uint32_t* d_data, size_t col_stride, int log_n, size_t cols, bool inverse) {
    if (log_n == 0 || cols == 0) return 0;
    SWIRL_REQUIRE(log_n <= 30, "log_n");
    const int tb = log_n < 11 ? log_n : 11;
    zeta_tile_kernel<<<(unsigned)(cols << (log_n - tb)), MLE_BLOCK, size_t(4) << tb, ctx->stream>>>(d_data, col_stride, log_n, tb,
                                                                                                 inverse ? 1 : 0);
    SWIRL_LAUNCH_CHECK(ctx);
    SWIRL_REQUIRE(log_n <= tb || ((col_stride & 3) == 0 && ((uintptr_t)d_data & 15) == 0), "zeta: alignment");
    for (int bit = tb; bit < log_n; bit++) {
        const size_t work = cols << (log_n - 3);
        zeta_stage_kernel<<<(unsigned)((work + MLE_BLOCK - 1) / MLE_BLOCK), MLE_BLOCK, 0, ctx->stream>>>(d_data, col_stride, log_n,
                                                                                                       cols, bit, inverse ? 1 : 0);
        SWIRL_LAUNCH_CHECK(ctx);
    }
    return 0;
}

// EF array-of-structs <-> 4 component columns
__global__ void __launch_bounds__(MLE_BLOCK)
ext_aos_to_soa_kernel(const uint32_t* __restrict__ aos, uint32_t* __restrict__ soa, size_t n, size_t col_stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Ext e = ldg_ext(aos + i * 4);
#pragma unroll
    for (int k = 0; k < 4; k++) soa[k * col_stride + i] = e.c[k];
}
__global__ void __launch_bounds__(MLE_BLOCK)
ext_soa_to_aos_kernel(const uint32_t* __restrict__ soa, uint32_t* __restrict__ aos, size_t n, size_t col_stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_ext(aos + i * 4, Ext{{soa[i], soa[col_stride + i], soa[2 * col_stride + i], soa[3 * col_stride + i]}});
}
int ext_aos_to_soa(swirl_ctx* ctx, const uint32_t* aos, uint32_t* soa, size_t n, size_t col_stride) {
    if (!n) return 0;
    ext_aos_to_soa_kernel<<<(unsigned)((n + MLE_BLOCK - 1) / MLE_BLOCK), MLE_BLOCK, 0, ctx->stream>>>(aos, soa, n, col_stride);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}
int ext_soa_to_aos(swirl_ctx* ctx, const uint32_t* soa, uint32_t* aos, size_t n, size_t col_stride) {
    if (!n) return 0;
    ext_soa_to_aos_kernel<<<(unsigned)((n + MLE_BLOCK - 1) / MLE_BLOCK), MLE_BLOCK, 0, ctx->stream>>>(soa, aos, n, col_stride);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}

// out[j] = in[2j] + (in[2j+1] - in[2j]) * r over a flat EF array: folds the low variable of every
// column of a column-major EF matrix with even height (fold_mle_evals, sumcheck.rs:395-414)
__global__ void __launch_bounds__(MLE_BLOCK)
ef_fold_flat_kernel2(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n_out, Ext r) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    st_ext(out + 4 * j, ext_lerp(ldg_ext(in + 8 * j), ldg_ext(in + 8 * j + 4), r));
}
int ef_fold_flat(swirl_ctx* ctx, const uint32_t* in, uint32_t* out, size_t n_out, const Ext& r) {
    if (!n_out) return 0;
    ef_fold_flat_kernel2<<<(unsigned)((n_out + MLE_BLOCK - 1) / MLE_BLOCK), MLE_BLOCK, 0, ctx->stream>>>(in, out, n_out, r);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}

// out[c][x] = sum_i L_i * mat[c][(x * 2^l + i + rot) mod height]: every 2^l_skip chunk's interpolant
// over D evaluated at the point whose Lagrange coefficients are L (fold_ple_evals, sumcheck.rs:204-251).
// out is column-major EF with height max(height, 2^l) >> l.
__global__ void __launch_bounds__(MLE_BLOCK)
fold_ple_kernel(const uint32_t* __restrict__ mat, size_t height, size_t new_height, size_t total, int l_skip, int rot,
                LagrangeArgs la, uint32_t* __restrict__ out) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const size_t c = g / new_height, x = g % new_height;
    const uint32_t* col = mat + c * height;
    Ext acc = bb::ext_zero();
    const int N = 1 << l_skip;
    int i = 0;
    for (; i + 4 <= N; i += 4) {  // four terms per Montgomery reduction (bb::dot4)
        uint32_t q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) q[j] = __ldg(col + ((x << l_skip) + i + j + rot) % height);
#pragma unroll
        for (int k = 0; k < 4; k++)
            acc.c[k] = bb::add(acc.c[k], bb::dot4(la.L[i][k], q[0], la.L[i + 1][k], q[1], la.L[i + 2][k], q[2], la.L[i + 3][k], q[3]));
    }
    for (; i < N; i++) {
        const size_t row = ((x << l_skip) + i + rot) % height;
        acc = bb::ext_add(acc, bb::ext_mul_base(Ext{{la.L[i][0], la.L[i][1], la.L[i][2], la.L[i][3]}}, __ldg(col + row)));
    }
    st_ext(out + 4 * g, acc);
}
int fold_ple(swirl_ctx* ctx, const uint32_t* mat, size_t height, size_t width, bool is_rot, int l_skip, const LagrangeArgs& la,
             uint32_t* out) {
    SWIRL_REQUIRE(l_skip <= 6, "l_skip > 6 unsupported");
    const size_t N = size_t(1) << l_skip;
    const size_t new_height = (height > N ? height : N) >> l_skip, total = new_height * width;
    if (!total) return 0;
    fold_ple_kernel<<<(unsigned)((total + MLE_BLOCK - 1) / MLE_BLOCK), MLE_BLOCK, 0, ctx->stream>>>(mat, height, new_height, total,
                                                                                              l_skip, is_rot ? 1 : 0, la, out);
    SWIRL_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace swirl

extern "C" int swirl_fold_mle(swirl_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, size_t n_out, const uint32_t r[4]) {
    SWIRL_REQUIRE(ctx && r && ((d_in && d_out) || n_out == 0), "null argument");
    SWIRL_REQUIRE((((uintptr_t)d_in | (uintptr_t)d_out) & 15) == 0, "EF buffers must be 16-byte aligned");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    return swirl::ef_fold_flat(ctx, d_in, d_out, n_out, bb::Ext{{r[0], r[1], r[2], r[3]}});
}
