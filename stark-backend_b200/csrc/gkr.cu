// LogUp-GKR fractional sumcheck: fraction tree + per-layer degree-3 sumchecks.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-backend/cuda/src/logup_zerocheck/gkr.cu              tree build, round, fold kernels
//   crates/cuda-backend/src/logup_zerocheck/fractional.rs:649-       fractional_sumcheck_gpu
// Semantics = crates/stark-backend/src/prover/logup_zerocheck/fractional_sumcheck_gkr.rs:60-213.
//
// Design: the tree is stored as in the reference CPU prover (node i has children 2i, 2i+1, layer k
// at [2^k, 2^(k+1)), Frac = {p, q} 32 bytes), leaves stay in the caller's buffer.  Layer j's
// sumcheck runs on 5 columns (eq, p0, q0, p1, q1).  Round 0 and round 1 read the tree layer
// directly (4 consecutive Fracs per hypercube point y = one 128-byte line) together with the eq
// table; from round 1 on every kernel FOLDS its input with the previous challenge on the fly,
// writes the folded table (half the size, structure-of-arrays) and accumulates the next round
// polynomial in the same pass, so each table is read once and written once per round.
// s(1), s(2), s(3) leave the device through one grid-wide reduction per round (ext.cuh) into mapped
// pinned memory; the host transcript (transcript.hpp) turns them into the next challenge.
#include <cstring>
#include <vector>

#include "ext.cuh"
#include "kernels.cuh"
#include "transcript.hpp"

namespace swirl {

using bb::ext_add;
using bb::ext_mul;
using bb::ext_sub;

constexpr int GKR_BLOCK = 256;

struct Frac4 {  // the four fractions tree[4y .. 4y+3] of one y, as rows x = 2y (lo) and 2y+1 (hi)
    Ext p0_lo, q0_lo, p1_lo, q1_lo, p0_hi, q0_hi, p1_hi, q1_hi;
};

// parent[i] = child[2i] + child[2i+1]  (projective fraction addition)
__global__ void __launch_bounds__(GKR_BLOCK)
frac_tree_layer_kernel(const uint32_t* __restrict__ child, uint32_t* __restrict__ parent, size_t n_parent) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_parent) return;
    const uint32_t* c = child + i * 16;
    const Ext pl = ldg_ext(c), ql = ldg_ext(c + 4), pr = ldg_ext(c + 8), qr = ldg_ext(c + 12);
    st_ext(parent + i * 8, ext_add(ext_mul(pl, qr), ext_mul(ql, pr)));
    st_ext(parent + i * 8 + 4, ext_mul(ql, qr));
}

struct XiArgs {
    uint32_t x[28][4];
};

// out[i] = prod_b (bit b of i ? x_b : 1 - x_b), i < 2^n  (poly.rs:133-149 evals_eq_hypercube).
// Two-level: `lo_bits` low variables come from a table built by the same kernel in a first launch.
__global__ void __launch_bounds__(GKR_BLOCK)
eq_table_kernel(XiArgs xi, int first_var, int n_vars, const uint32_t* __restrict__ lo_table, int lo_bits,
                uint32_t* __restrict__ out, size_t n_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    Ext acc = lo_table ? ldg_ext(lo_table + (i & ((size_t(1) << lo_bits) - 1)) * 4) : bb::ext_one();
    const size_t hi = lo_table ? i >> lo_bits : i;
    for (int b = 0; b < n_vars; b++) {
        const Ext x = Ext{{xi.x[first_var + b][0], xi.x[first_var + b][1], xi.x[first_var + b][2], xi.x[first_var + b][3]}};
        acc = ext_mul(acc, ((hi >> b) & 1) ? x : ext_one_minus(x));
    }
    st_ext(out + i * 4, acc);
}

// out[i] = A[i mod 2^lo_bits] * B[i >> lo_bits]: eq table of (x_lo, x_hi) from the two half tables
__global__ void __launch_bounds__(GKR_BLOCK)
eq_combine_kernel(const uint32_t* __restrict__ A, const uint32_t* __restrict__ B, int lo_bits,
                  uint32_t* __restrict__ out, size_t n_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    st_ext(out + i * 4, ext_mul(ldg_ext(A + (i & ((size_t(1) << lo_bits) - 1)) * 4), ldg_ext(B + (i >> lo_bits) * 4)));
}

struct RoundArgs {
    const uint32_t* tree;  // FROM_TREE: layer segment, 2 * height Fracs
    const uint32_t* eq;    // FROM_TREE: eq table, `height` EF
    const uint32_t* in;    // !FROM_TREE: 5 columns (eq, p0, q0, p1, q1) of `height` EF, column stride in_stride
    size_t in_stride;
    uint32_t* out;  // FOLD: 5 columns of height/2 EF, column stride out_stride
    size_t out_stride;
    size_t height;  // rows of the input table
    uint32_t r[4];  // FOLD: previous challenge
    uint32_t lambda[4];
    uint32_t* partials;
    unsigned int* ticket;
    uint32_t* result;  // 12 words: s(1), s(2), s(3)
};

template <bool FROM_TREE>
__device__ __forceinline__ void load_row(const RoundArgs& a, size_t x, Ext (&row)[5]) {
    if (FROM_TREE) {
        row[0] = ldg_ext(a.eq + x * 4);
        const uint32_t* t = a.tree + x * 16;
        row[1] = ldg_ext(t);
        row[2] = ldg_ext(t + 4);
        row[3] = ldg_ext(t + 8);
        row[4] = ldg_ext(t + 12);
    } else {
#pragma unroll
        for (int c = 0; c < 5; c++) row[c] = ldg_ext(a.in + (c * a.in_stride + x) * 4);
    }
}

// contribution of the pair (lo, hi) = rows (2y, 2y+1) to s(1), s(2), s(3)
__device__ __forceinline__ void accumulate(const Ext (&lo)[5], const Ext (&hi)[5], const Ext& lambda, Ext (&s)[3]) {
    Ext d[5], cur[5];
#pragma unroll
    for (int c = 0; c < 5; c++) {
        d[c] = ext_sub(hi[c], lo[c]);
        cur[c] = hi[c];  // X = 1
    }
#pragma unroll
    for (int X = 0; X < 3; X++) {
        // eq * (p0 q1 + p1 q0 + lambda q0 q1) = eq * (p0 q1 + q0 (p1 + lambda q1))
        const Ext inner = ext_add(ext_mul(cur[1], cur[4]), ext_mul(cur[2], ext_add(cur[3], ext_mul(lambda, cur[4]))));
        s[X] = ext_add(s[X], ext_mul(cur[0], inner));
        if (X < 2) {
#pragma unroll
            for (int c = 0; c < 5; c++) cur[c] = ext_add(cur[c], d[c]);
        }
    }
}

// One sumcheck round.  FOLD = false: table rows are read as they are (first round of a layer).
// FOLD = true: rows 4y..4y+3 are folded pairwise with a.r into the two rows 2y, 2y+1 of the next
// table, which are written out and used for this round's polynomial.
template <bool FROM_TREE, bool FOLD>
__global__ void __launch_bounds__(GKR_BLOCK) gkr_round_kernel(RoundArgs a) {
    const Ext lambda = Ext{{a.lambda[0], a.lambda[1], a.lambda[2], a.lambda[3]}};
    const Ext r = Ext{{a.r[0], a.r[1], a.r[2], a.r[3]}};
    Ext s[3] = {bb::ext_zero(), bb::ext_zero(), bb::ext_zero()};
    const size_t ny = FOLD ? a.height >> 2 : a.height >> 1;
    for (size_t y = (size_t)blockIdx.x * blockDim.x + threadIdx.x; y < ny; y += (size_t)gridDim.x * blockDim.x) {
        Ext lo[5], hi[5];
        if (FOLD) {
            Ext a0[5], a1[5];
            load_row<FROM_TREE>(a, 4 * y, a0);
            load_row<FROM_TREE>(a, 4 * y + 1, a1);
#pragma unroll
            for (int c = 0; c < 5; c++) lo[c] = ext_lerp(a0[c], a1[c], r);
            load_row<FROM_TREE>(a, 4 * y + 2, a0);
            load_row<FROM_TREE>(a, 4 * y + 3, a1);
#pragma unroll
            for (int c = 0; c < 5; c++) hi[c] = ext_lerp(a0[c], a1[c], r);
#pragma unroll
            for (int c = 0; c < 5; c++) {
                st_ext(a.out + (c * a.out_stride + 2 * y) * 4, lo[c]);
                st_ext(a.out + (c * a.out_stride + 2 * y + 1) * 4, hi[c]);
            }
        } else {
            load_row<FROM_TREE>(a, 2 * y, lo);
            load_row<FROM_TREE>(a, 2 * y + 1, hi);
        }
        accumulate(lo, hi, lambda, s);
    }
    uint32_t v[12];
#pragma unroll
    for (int X = 0; X < 3; X++)
#pragma unroll
        for (int k = 0; k < 4; k++) v[X * 4 + k] = s[X].c[k];
    grid_sum<12>(v, a.partials, a.ticket, a.result);
}

// Last fold of a layer: the 2-row table -> its single row = the four layer claims (p0, q0, p1, q1).
template <bool FROM_TREE>
__global__ void gkr_claims_kernel(RoundArgs a) {
    if (threadIdx.x >= 4) return;
    const Ext r = Ext{{a.r[0], a.r[1], a.r[2], a.r[3]}};
    Ext lo[5], hi[5];
    load_row<FROM_TREE>(a, 0, lo);
    load_row<FROM_TREE>(a, 1, hi);
    Ext sel_lo = lo[1], sel_hi = hi[1];
#pragma unroll
    for (int c = 2; c < 5; c++)
        if ((int)threadIdx.x == c - 1) {
            sel_lo = lo[c];
            sel_hi = hi[c];
        }
    const Ext v = ext_lerp(sel_lo, sel_hi, r);
#pragma unroll
    for (int k = 0; k < 4; k++) a.result[threadIdx.x * 4 + k] = v.c[k];
}

static int round_grid(const swirl_ctx* ctx, size_t work_items) {
    size_t blocks = (work_items + GKR_BLOCK - 1) / GKR_BLOCK;
    const size_t cap = (size_t)ctx->sm_count * 4;
    if (blocks > cap) blocks = cap;
    return blocks ? (int)blocks : 1;
}

static Ext ext_from_words(const uint32_t* w) { return Ext{{w[0], w[1], w[2], w[3]}}; }

}  // namespace swirl

using namespace swirl;

extern "C" int swirl_gkr_fractional_sumcheck(swirl_ctx* ctx, swirl_transcript* ts, const uint32_t* d_leaves, int log_n,
                                             int assert_zero, uint32_t h_frac_sum[8], uint32_t* h_claims,
                                             uint32_t* h_polys, uint32_t* h_xi) {
    SWIRL_REQUIRE(ctx && ts && d_leaves && h_frac_sum && h_claims && h_xi, "null argument");
    SWIRL_REQUIRE(log_n >= 1 && log_n <= 27, "log_n must be in [1, 27]");
    SWIRL_REQUIRE(((uintptr_t)d_leaves & 15) == 0, "leaves must be 16-byte aligned");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    Transcript tr(ts);
    RoundScratch* rs;
    SWIRL_TRY(round_scratch_get(ctx, &rs));
    const int n = log_n;
    const size_t N = size_t(1) << n;

    // ---- tree: layers 0..n-1 in `tree` (node i at tree + i*8 words), layer n = the leaves ------
    uint32_t* tree = nullptr;
    SWIRL_CUDA(dev_alloc(ctx, &tree, N * 8));
    auto layer_ptr = [&](int k) -> const uint32_t* { return k == n ? d_leaves : tree + (size_t(1) << k) * 8; };
    for (int k = n - 1; k >= 0; k--) {
        const size_t np = size_t(1) << k;
        frac_tree_layer_kernel<<<(unsigned)((np + GKR_BLOCK - 1) / GKR_BLOCK), GKR_BLOCK, 0, ctx->stream>>>(
            layer_ptr(k + 1), tree + np * 8, np);
        SWIRL_LAUNCH_CHECK(ctx);
    }
    // root + layer 1 (nodes 1, 2, 3) -> host
    uint32_t top[24];
    SWIRL_CUDA(cudaMemcpyAsync(top, tree + 8, n >= 2 ? 96 : 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (n == 1) SWIRL_CUDA(cudaMemcpyAsync(top + 8, d_leaves, 64, cudaMemcpyDeviceToHost, ctx->stream));
    SWIRL_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(h_frac_sum, top, 32);
    int rc = 0;
    const Ext root_p = ext_from_words(top), root_q = ext_from_words(top + 4);
    if (assert_zero) {
        if (root_p.c[0] | root_p.c[1] | root_p.c[2] | root_p.c[3]) {
            set_error("LogupZerocheckError::NonZeroRootSum");
            dev_free(ctx, tree);
            return SWIRL_ERR_NONZERO_ROOT_SUM;
        }
    } else {
        tr.observe_ext(root_p);
    }
    tr.observe_ext(root_q);
    // layer 1 claims: tree[2].p, tree[2].q, tree[3].p, tree[3].q
    memcpy(h_claims, top + 8, 64);
    for (int i = 0; i < 4; i++) tr.observe_ext(ext_from_words(top + 8 + 4 * i));
    std::vector<Ext> xi_prev{tr.sample_ext()};

    // ---- working buffers: eq table (<= N/2 EF), two ping-pong SoA tables (5 cols x <= N/4 EF) ---
    uint32_t *eq = nullptr, *eq_lo = nullptr, *tab[2] = {nullptr, nullptr};
    const size_t max_h = N >> 1;  // largest layer table height (layer n-1)
    const size_t tab_stride = max_h >= 2 ? max_h >> 1 : 1;
    if (n >= 2) {
        SWIRL_CUDA(dev_alloc(ctx, &eq, max_h * 4));
        SWIRL_CUDA(dev_alloc(ctx, &eq_lo, (size_t(8) << 14)));  // two half tables of <= 2^14 EF
        SWIRL_CUDA(dev_alloc(ctx, &tab[0], 5 * tab_stride * 4));
        SWIRL_CUDA(dev_alloc(ctx, &tab[1], 5 * (tab_stride >> 1 ? tab_stride >> 1 : 1) * 4));
    }
    size_t poly_off = 0;
    for (int round = 1; round < n && rc == 0; round++) {
        const size_t H = size_t(1) << round;
        const Ext lambda = tr.sample_ext();
        // eq table of xi_prev (round variables)
        XiArgs xa;
        for (int b = 0; b < round; b++)
            for (int k = 0; k < 4; k++) xa.x[b][k] = xi_prev[b].c[k];
        if (round <= 12) {
            eq_table_kernel<<<(unsigned)((H + GKR_BLOCK - 1) / GKR_BLOCK), GKR_BLOCK, 0, ctx->stream>>>(xa, 0, round, nullptr, 0,
                                                                                                      eq, H);
        } else {
            const int lo_bits = round / 2, hi_bits = round - lo_bits;  // both <= 14
            uint32_t* eq_hi = eq_lo + (size_t(4) << 14);
            eq_table_kernel<<<(unsigned)(((size_t(1) << lo_bits) + GKR_BLOCK - 1) / GKR_BLOCK), GKR_BLOCK, 0, ctx->stream>>>(
                xa, 0, lo_bits, nullptr, 0, eq_lo, size_t(1) << lo_bits);
            eq_table_kernel<<<(unsigned)(((size_t(1) << hi_bits) + GKR_BLOCK - 1) / GKR_BLOCK), GKR_BLOCK, 0, ctx->stream>>>(
                xa, lo_bits, hi_bits, nullptr, 0, eq_hi, size_t(1) << hi_bits);
            ctx->launches += 2;
            eq_combine_kernel<<<(unsigned)((H + GKR_BLOCK - 1) / GKR_BLOCK), GKR_BLOCK, 0, ctx->stream>>>(eq_lo, eq_hi, lo_bits, eq, H);
        }
        SWIRL_LAUNCH_CHECK(ctx);

        RoundArgs a{};
        a.tree = layer_ptr(round + 1);
        a.eq = eq;
        a.partials = rs->d_partials;
        a.ticket = rs->d_ticket;
        a.result = rs->d_result;
        for (int k = 0; k < 4; k++) a.lambda[k] = lambda.c[k];
        std::vector<Ext> rho;
        size_t h = H;  // height of the table the next kernel reads
        int cur = 0;
        for (int sr = 0; sr < round; sr++) {
            SwirlTimed timed(ctx, SWIRL_T_GKR);
            if (sr == 0) {
                a.height = H;
                gkr_round_kernel<true, false><<<round_grid(ctx, H >> 1), GKR_BLOCK, 0, ctx->stream>>>(a);
            } else if (sr == 1) {
                a.height = H;
                a.out = tab[0];
                a.out_stride = tab_stride;
                gkr_round_kernel<true, true><<<round_grid(ctx, H >> 2), GKR_BLOCK, 0, ctx->stream>>>(a);
                h = H >> 1;
                cur = 0;
            } else {
                a.in = tab[cur];
                a.in_stride = cur == 0 ? tab_stride : (tab_stride >> 1 ? tab_stride >> 1 : 1);
                a.height = h;
                a.out = tab[cur ^ 1];
                a.out_stride = cur == 0 ? (tab_stride >> 1 ? tab_stride >> 1 : 1) : tab_stride;
                gkr_round_kernel<false, true><<<round_grid(ctx, h >> 2), GKR_BLOCK, 0, ctx->stream>>>(a);
                h >>= 1;
                cur ^= 1;
            }
            SWIRL_LAUNCH_CHECK(ctx);
            SWIRL_CUDA(cudaStreamSynchronize(ctx->stream));
            uint32_t* out = h_polys + (poly_off + sr) * 12;
            memcpy(out, rs->h_result, 48);
            for (int X = 0; X < 3; X++) tr.observe_ext(ext_from_words(out + 4 * X));
            const Ext r = tr.sample_ext();
            rho.push_back(r);
            for (int k = 0; k < 4; k++) a.r[k] = r.c[k];
        }
        // claims: fold the remaining 2-row table with the last challenge
        if (round == 1) {
            a.height = 2;
            gkr_claims_kernel<true><<<1, 32, 0, ctx->stream>>>(a);
        } else {
            a.in = tab[cur];
            a.in_stride = cur == 0 ? tab_stride : (tab_stride >> 1 ? tab_stride >> 1 : 1);
            a.height = 2;
            gkr_claims_kernel<false><<<1, 32, 0, ctx->stream>>>(a);
        }
        SWIRL_LAUNCH_CHECK(ctx);
        SWIRL_CUDA(cudaStreamSynchronize(ctx->stream));
        uint32_t* cl = h_claims + (size_t)round * 16;
        memcpy(cl, rs->h_result, 64);
        for (int i = 0; i < 4; i++) tr.observe_ext(ext_from_words(cl + 4 * i));
        const Ext mu = tr.sample_ext();
        xi_prev.assign(1, mu);
        xi_prev.insert(xi_prev.end(), rho.begin(), rho.end());
        poly_off += round;
    }
    for (int b = 0; b < n; b++) memcpy(h_xi + 4 * b, xi_prev[b].c, 16);
    dev_free(ctx, tree);
    dev_free(ctx, eq);
    dev_free(ctx, eq_lo);
    dev_free(ctx, tab[0]);
    dev_free(ctx, tab[1]);
    return rc;
}
