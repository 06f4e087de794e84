// LogUp-GKR fractional sumcheck: fraction tree + per-layer degree-3 sumchecks.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-backend/cuda/src/logup_zerocheck/gkr.cu              tree build, round, fold kernels
//   crates/cuda-backend/src/logup_zerocheck/fractional.rs:649-       fractional_sumcheck_gpu
// Semantics = crates/stark-backend/src/prover/logup_zerocheck/fractional_sumcheck_gkr.rs:60-213.
//
// Design: the tree is stored as in the reference CPU prover (node i has children 2i, 2i+1, Frac =
// {p, q} 32 bytes), leaves stay in the caller's buffer.  Three things keep the work proportional to
// the REAL interactions and the traffic close to one read per table per round:
//  * constant tail: the interaction layout pads the leaf layer with (0, alpha) up to the next power
//    of two (between 0 and 50 % of it).  Every node above a constant run is again a constant
//    (0, c_k), c_k = c_{k+1}^2, so each layer stores only its first S_k nodes and the sumchecks run
//    over the stored rows; the tail's contribution lambda * c^2 * sum_{y >= y_tail} eq(xi, y) has a
//    closed form in the O(log) prefix sums of eq and is added on the host (field arithmetic is
//    exact, so the round polynomials are the same elements the full-table sum would give).
//  * eq is never materialised: eq(xi,(x_0..x_{s-1} = r, X, y)) = e_bound * eq1(xi_s, X) * E_s[y],
//    the first two factors leave the sum (host), E_s[y] = A_s[y_lo] * B_s[y_hi] comes from suffix
//    tables of <= 2^14 entries that stay in L2 (eq_suffix_kernel).
//  * layer j's sumcheck runs on 4 columns (p0, q0, p1, q1).  Round 0 and round 1 read the tree layer
//    directly (4 consecutive Fracs per pair = one 128-byte line); from round 1 on every kernel FOLDS
//    its input with the previous challenge on the fly, writes the folded table (half the size,
//    structure-of-arrays) and accumulates the next round polynomial in the same pass.
// t(1), t(2), t(3) (the eq-free inner sums) leave the device through one grid-wide reduction per
// round (ext.cuh) into mapped pinned memory; the host transcript (transcript.hpp) turns them into
// s(1), s(2), s(3) and the next challenge.
#include <array>
#include <cstring>
#include <vector>

#include "ext.cuh"
#include "hostpoly.hpp"
#include "kernels.cuh"
#include "transcript.hpp"

namespace swirl {

using bb::ext_add;
using bb::ext_mul;
using bb::ext_mul_base;
using bb::ext_sub;

constexpr int GKR_BLOCK = 256;
constexpr int GKR_HOST_LOG = 5;  // sumcheck rounds on tables of <= 2^4 rows run on the host (a kernel round trip costs more)

// parent[i] = child[2i] + child[2i+1] (projective fraction addition) for the stored prefix of a layer;
// parents whose children lie in the constant tail are the constant (0, c_parent).
__global__ void __launch_bounds__(GKR_BLOCK)
frac_tree_layer_kernel(const uint32_t* __restrict__ child, uint32_t* __restrict__ parent, size_t n_parent,
                       size_t n_child, Ext c_parent) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_parent) return;
    if (2 * i >= n_child) {
        st_ext(parent + i * 8, bb::ext_zero());
        st_ext(parent + i * 8 + 4, c_parent);
        return;
    }
    const uint32_t* c = child + i * 16;
    const Ext pl = ldg_ext(c), ql = ldg_ext(c + 4), pr = ldg_ext(c + 8), qr = ldg_ext(c + 12);
    st_ext(parent + i * 8, ext_add(ext_mul(pl, qr), ext_mul(ql, pr)));
    st_ext(parent + i * 8 + 4, ext_mul(ql, qr));
}

struct XiArgs {
    uint32_t x[28][4];
};

// Suffix eq tables of the variables [first_var, first_var + n_vars): table t (blockIdx.y) covers the
// variables [first_var + t, first_var + n_vars), has 2^(n_vars - t) entries
// out_t[i] = prod_b (bit b of i ? x_{first_var+t+b} : 1 - x_{first_var+t+b})   (poly.rs:133-149)
// and starts at entry 2^(n_vars+1) - 2^(n_vars-t+1); table n_vars is the single entry 1.
__global__ void __launch_bounds__(GKR_BLOCK) eq_suffix_kernel(XiArgs xi, int first_var, int n_vars, uint32_t* __restrict__ out) {
    const int t = blockIdx.y;
    const size_t size = size_t(1) << (n_vars - t);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    Ext acc = bb::ext_one();
    for (int b = 0; b < n_vars - t; b++) {
        const uint32_t* xw = xi.x[first_var + t + b];
        const Ext x = Ext{{xw[0], xw[1], xw[2], xw[3]}};
        acc = ext_mul(acc, ((i >> b) & 1) ? x : ext_one_minus(x));
    }
    st_ext(out + ((size_t(2) << n_vars) - 2 * size + i) * 4, acc);
}

struct RoundArgs {
    const uint32_t* tree;  // FROM_TREE: layer segment; row x = Fracs 2x, 2x+1
    const uint32_t* in;    // !FROM_TREE: 4 columns (p0, q0, p1, q1) of EF, column stride in_stride
    size_t in_stride;
    size_t rows_in;  // stored rows of the input table; rows beyond are the constant (0, c, 0, c)
    uint32_t* out;   // FOLD: 4 columns, column stride out_stride
    size_t out_stride;
    size_t ny;          // pairs (FOLD: quads) to process
    const uint32_t* A;  // eq suffix table of the low variables (a_bits of them), or unused when a_bits == 0
    const uint32_t* B;  // eq suffix table of the high variables
    int a_bits;
    int g_bits;     // != 0: runs of 2^g_bits pairs (g_bits <= a_bits) per block, see gkr_round_kernel
    uint32_t c[4];  // the constant q of the tail rows
    uint32_t r[4];  // FOLD: previous challenge
    uint32_t lambda[4];
    uint32_t* partials;
    unsigned int* ticket;
    uint32_t* result;  // 8 words: t(1) and the leading coefficient of t
    uint32_t* last_rows;  // non-null in the last device round of a layer: this round's whole table (rows 2y = lo, 2y+1 = hi,
                          // 4 EF each, at most 32 rows) for the host, which finishes the layer
    RoundLink link;       // seq != 0: the previous challenge arrives through the mailbox instead of `r` (ext.cuh)
};

template <bool FROM_TREE>
__device__ __forceinline__ void load_row(const RoundArgs& a, size_t x, const Ext& c, Ext (&row)[4]) {
    if (x >= a.rows_in) {
        row[0] = bb::ext_zero();
        row[1] = c;
        row[2] = bb::ext_zero();
        row[3] = c;
        return;
    }
    if (FROM_TREE) {
        const uint32_t* t = a.tree + x * 16;
#pragma unroll
        for (int k = 0; k < 4; k++) row[k] = ldg_ext(t + 4 * k);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) row[k] = ldg_ext(a.in + (k * a.in_stride + x) * 4);
    }
}

// contribution of the pair (lo, hi) = rows (2y, 2y+1) to t(1) and to the leading coefficient of the quadratic
// t(X) = sum_y E[y] * inner(X, y), inner = p0 q1 + p1 q0 + lambda q0 q1 = p0 q1 + q0 (p1 + lambda q1) with every
// factor linear in X.  Two values are enough: t(0) follows on the host from the sumcheck identity s(0) + s(1) = claim,
// so the round costs 8 extension multiplications per pair instead of 12.
__host__ __device__ __forceinline__ void accumulate(const Ext (&lo)[4], const Ext (&hi)[4], const Ext& lambda, const Ext& E, Ext (&s)[2]) {
    const Ext d0 = ext_sub(hi[0], lo[0]), d1 = ext_sub(hi[1], lo[1]), d2 = ext_sub(hi[2], lo[2]), d3 = ext_sub(hi[3], lo[3]);
    const Ext w_hi = ext_add(hi[2], ext_mul(lambda, hi[3])), dw = ext_add(d2, ext_mul(lambda, d3));
    const Ext at1 = ext_add(ext_mul(hi[0], hi[3]), ext_mul(hi[1], w_hi));
    const Ext lead = ext_add(ext_mul(d0, d3), ext_mul(d1, dw));
    s[0] = ext_add(s[0], ext_mul(E, at1));
    s[1] = ext_add(s[1], ext_mul(E, lead));
}

// One sumcheck round.  FOLD = false: table rows are read as they are (first round of a layer).
// FOLD = true: rows 4y..4y+3 are folded pairwise with a.r into the two rows 2y, 2y+1 of the next
// table, which are written out and used for this round's polynomial.
template <bool FROM_TREE, bool FOLD, bool RUNS>
__global__ void __launch_bounds__(GKR_BLOCK, (FOLD || RUNS) ? 3 : 4) gkr_round_kernel(RoundArgs a) {
    const Ext lambda = Ext{{a.lambda[0], a.lambda[1], a.lambda[2], a.lambda[3]}};
    Ext r = Ext{{a.r[0], a.r[1], a.r[2], a.r[3]}};
    if (FOLD && !link_wait(a.link, r)) return;
    const Ext c = Ext{{a.c[0], a.c[1], a.c[2], a.c[3]}};
    Ext s[2] = {bb::ext_zero(), bb::ext_zero()};
    const size_t a_mask = (size_t(1) << a.a_bits) - 1;
    // rows of the pair y -> (lo, hi), folded / written / exported as the round requires, and their share of the sums with weight E
    auto pair = [&](size_t y, const Ext& E, Ext (&acc)[2]) {
        Ext lo[4], hi[4];
        if (FOLD) {
            Ext a0[4], a1[4];
            load_row<FROM_TREE>(a, 4 * y, c, a0);
            load_row<FROM_TREE>(a, 4 * y + 1, c, a1);
#pragma unroll
            for (int k = 0; k < 4; k++) lo[k] = ext_lerp(a0[k], a1[k], r);
            load_row<FROM_TREE>(a, 4 * y + 2, c, a0);
            load_row<FROM_TREE>(a, 4 * y + 3, c, a1);
#pragma unroll
            for (int k = 0; k < 4; k++) hi[k] = ext_lerp(a0[k], a1[k], r);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                st_ext(a.out + (k * a.out_stride + 2 * y) * 4, lo[k]);
                st_ext(a.out + (k * a.out_stride + 2 * y + 1) * 4, hi[k]);
            }
        } else {
            load_row<FROM_TREE>(a, 2 * y, c, lo);
            load_row<FROM_TREE>(a, 2 * y + 1, c, hi);
        }
        if (a.last_rows) {  // the host finishes the layer from this (small) table
#pragma unroll
            for (int k = 0; k < 4; k++) {
                st_ext(a.last_rows + (2 * y) * 16 + 4 * k, lo[k]);
                st_ext(a.last_rows + (2 * y + 1) * 16 + 4 * k, hi[k]);
            }
        }
        accumulate(lo, hi, lambda, E, acc);
    };
    if (RUNS) {
        // Large tables: a block walks runs of 2^g_bits consecutive pairs, which share the high eq factor B[y >> a_bits]; the
        // run is summed with the low factor A only and multiplied by B once per thread (8 instead of 9 EF products per pair).
        const size_t run = size_t(1) << a.g_bits, n_runs = (a.ny + run - 1) >> a.g_bits;
        for (size_t u = blockIdx.x; u < n_runs; u += gridDim.x) {
            const size_t y0 = u << a.g_bits, yl0 = y0 & a_mask;
            Ext t[2] = {bb::ext_zero(), bb::ext_zero()};
            for (size_t yl = threadIdx.x; yl < run && y0 + yl < a.ny; yl += blockDim.x) pair(y0 + yl, ldg_ext(a.A + (yl0 + yl) * 4), t);
            const Ext Bv = ldg_ext(a.B + (y0 >> a.a_bits) * 4);
            s[0] = ext_add(s[0], ext_mul(Bv, t[0]));
            s[1] = ext_add(s[1], ext_mul(Bv, t[1]));
        }
    } else {
        for (size_t y = (size_t)blockIdx.x * blockDim.x + threadIdx.x; y < a.ny; y += (size_t)gridDim.x * blockDim.x) {
            Ext E = ldg_ext(a.B + (y >> a.a_bits) * 4);
            if (a.a_bits) E = ext_mul(E, ldg_ext(a.A + (y & a_mask) * 4));
            pair(y, E, s);
        }
    }
    uint32_t v[8];
#pragma unroll
    for (int X = 0; X < 2; X++)
#pragma unroll
        for (int k = 0; k < 4; k++) v[X * 4 + k] = s[X].c[k];
    if (a.last_rows) __threadfence_system();  // the exported table must be in host memory before the result words say "ready"
    grid_sum<8>(v, a.partials, a.ticket, a.result, link_result_tag(a.link.seq));
}

// `per_sm` = resident blocks per SM of the variant launched (its __launch_bounds__): a grid-stride sweep over more
// blocks than fit at once runs a second, mostly empty wave
static int round_grid(const swirl_ctx* ctx, size_t work_items, int per_sm = 4) {
    size_t blocks = (work_items + GKR_BLOCK - 1) / GKR_BLOCK;
    const size_t cap = (size_t)ctx->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    return blocks ? (int)blocks : 1;
}

// sum_{y < t} prod_{i < count} eq1(xi[first + i], bit i of y), t <= 2^count
static Ext eq_prefix_sum(const std::vector<Ext>& xi, int first, int count, size_t t) {
    if (t >= (size_t(1) << count)) return bb::ext_one();
    Ext acc = bb::ext_zero(), pre = bb::ext_one();
    for (int i = count - 1; i >= 0; i--) {
        const Ext& x = xi[first + i];
        if ((t >> i) & 1) {
            acc = ext_add(acc, ext_mul(pre, ext_one_minus(x)));
            pre = ext_mul(pre, x);
        } else {
            pre = ext_mul(pre, ext_one_minus(x));
        }
    }
    return acc;
}

static Ext ext_from_words(const uint32_t* w) { return Ext{{w[0], w[1], w[2], w[3]}}; }

}  // namespace swirl

using namespace swirl;

static size_t round_up4(size_t x) { return (x + 3) & ~size_t(3); }

extern "C" int swirl_gkr_fractional_sumcheck_padded(swirl_ctx* ctx, swirl_transcript* ts, const uint32_t* d_leaves,
                                                    uint64_t n_stored, const uint32_t pad_q[4], int log_n, int assert_zero,
                                                    uint32_t h_frac_sum[8], uint32_t* h_claims, uint32_t* h_polys,
                                                    uint32_t* h_xi) {
    SWIRL_REQUIRE(ctx && ts && d_leaves && h_frac_sum && h_claims && h_xi, "null argument");
    SWIRL_REQUIRE(log_n >= 1 && log_n <= 27, "log_n must be in [1, 27]");
    SWIRL_REQUIRE(((uintptr_t)d_leaves & 15) == 0, "leaves must be 16-byte aligned");
    const int n = log_n;
    const size_t N = size_t(1) << n;
    SWIRL_REQUIRE(n_stored >= 1 && n_stored <= N && (n_stored == N || (n_stored % 4 == 0 && pad_q)),
                  "n_stored must be 2^log_n or a multiple of 4 with a padding denominator");
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    Transcript tr(ts);
    RoundScratch* rs;
    SWIRL_TRY(round_scratch_get(ctx, &rs));

    // ---- stored prefix S[k] and tail constant c[k] of every layer (layer n = the leaves) ----------
    std::vector<size_t> S(n + 1), off(n + 1, 0);
    std::vector<Ext> cst(n + 1, bb::ext_one());
    S[n] = n_stored;
    if (pad_q) cst[n] = ext_from_words(pad_q);
    size_t tree_nodes = 0;
    for (int k = n - 1; k >= 0; k--) {
        S[k] = std::min(size_t(1) << k, round_up4((S[k + 1] + 1) / 2));
        cst[k] = ext_mul(cst[k + 1], cst[k + 1]);
        off[k] = tree_nodes;
        tree_nodes += round_up4(S[k]);  // keeps every layer segment 128-byte aligned
    }
    uint32_t* tree = nullptr;
    ArenaGuard scratch(ctx);  // tree, eq tables, working tables: released on every return path
    SWIRL_CUDA(dev_alloc(ctx, &tree, tree_nodes * 8));
    scratch.add(tree);
    auto layer_ptr = [&](int k) -> const uint32_t* { return k == n ? d_leaves : tree + off[k] * 8; };
    for (int k = n - 1; k >= 0; k--) {
        frac_tree_layer_kernel<<<(unsigned)((S[k] + GKR_BLOCK - 1) / GKR_BLOCK), GKR_BLOCK, 0, ctx->stream>>>(
            layer_ptr(k + 1), tree + off[k] * 8, S[k], S[k + 1], cst[k]);
        SWIRL_LAUNCH_CHECK(ctx);
    }
    // root (layer 0) + layer 1 -> host, and the stored nodes of the layers whose sumchecks run entirely on the host
    // (tables of at most HOST_ROWS rows)
    uint32_t top[24];
    SWIRL_CUDA(cudaMemcpyAsync(top, layer_ptr(0), 32, cudaMemcpyDeviceToHost, ctx->stream));
    SWIRL_CUDA(cudaMemcpyAsync(top + 8, layer_ptr(1), 64, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<std::vector<uint32_t>> top_layers(std::min(n, GKR_HOST_LOG + 1) + 1);
    for (int k = 2; k < (int)top_layers.size(); k++) {
        top_layers[k].resize(S[k] * 8);
        SWIRL_CUDA(cudaMemcpyAsync(top_layers[k].data(), layer_ptr(k), S[k] * 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
    memcpy(h_frac_sum, top, 32);
    int rc = 0;
    const Ext root_p = ext_from_words(top), root_q = ext_from_words(top + 4);
    if (assert_zero) {
        if (root_p.c[0] | root_p.c[1] | root_p.c[2] | root_p.c[3]) {
            set_error("LogupZerocheckError::NonZeroRootSum");
            return SWIRL_ERR_NONZERO_ROOT_SUM;
        }
    } else {
        tr.observe_ext(root_p);
    }
    tr.observe_ext(root_q);
    // layer 1 claims: node 0 and node 1 of layer 1
    memcpy(h_claims, top + 8, 64);
    for (int i = 0; i < 4; i++) tr.observe_ext(ext_from_words(top + 8 + 4 * i));
    std::vector<Ext> xi_prev{tr.sample_ext()};

    // ---- working buffers: eq suffix tables, two ping-pong SoA tables (4 columns) -----------------
    uint32_t *eqA = nullptr, *eqB = nullptr, *tab[2] = {nullptr, nullptr};
    size_t tab_stride[2] = {1, 1};
    if (n >= 2) {
        SWIRL_CUDA(dev_alloc(ctx, &eqA, (size_t(8) << 14)));  // suffix tables of <= 13 variables: < 2^14 EF
        scratch.add(eqA);
        SWIRL_CUDA(dev_alloc(ctx, &eqB, (size_t(8) << 14)));
        scratch.add(eqB);
        const size_t rows_max = (S[n] + 1) / 2;                // stored rows of the largest layer table
        tab_stride[0] = 2 * ((rows_max + 3) / 4);              // rows written by the first fold
        tab_stride[1] = 2 * ((tab_stride[0] + 3) / 4);
        SWIRL_CUDA(dev_alloc(ctx, &tab[0], 4 * tab_stride[0] * 4));
        scratch.add(tab[0]);
        SWIRL_CUDA(dev_alloc(ctx, &tab[1], 4 * tab_stride[1] * 4));
        scratch.add(tab[1]);
    }
    const Ext one = bb::ext_one();
    const uint32_t inv2 = bb::inv(bb::mont(2)), inv6 = bb::inv(bb::mont(6));
    // claims of the previous layer: p(xi, 0), q(xi, 0), p(xi, 1), q(xi, 1)
    Ext prev_claims[4] = {ext_from_words(top + 8), ext_from_words(top + 12), ext_from_words(top + 16), ext_from_words(top + 20)};
    size_t poly_off = 0;
    for (int round = 1; round < n && rc == 0; round++) {
        const Ext lambda = tr.sample_ext();
        // eq(xi_prev, x) = e_bound * eq1(xi_s, X) * A_s[y_lo] * B_s[y_hi]: variable 0 never enters a table;
        // low group = variables [1, v_split), high group = [v_split, round)
        const int v_split = round <= 13 ? round : (round + 1) / 2 + 1;
        const int nA = v_split - 1, nB = round - v_split;  // both <= 13
        XiArgs xa;
        for (int b = 0; b < round; b++)
            for (int k = 0; k < 4; k++) xa.x[b][k] = xi_prev[b].c[k];
        {
            const dim3 gA((unsigned)(((size_t(1) << nA) + GKR_BLOCK - 1) / GKR_BLOCK), nA + 1);
            eq_suffix_kernel<<<gA, GKR_BLOCK, 0, ctx->stream>>>(xa, 1, nA, eqA);
            SWIRL_LAUNCH_CHECK(ctx);
            const dim3 gB((unsigned)(((size_t(1) << nB) + GKR_BLOCK - 1) / GKR_BLOCK), nB + 1);
            eq_suffix_kernel<<<gB, GKR_BLOCK, 0, ctx->stream>>>(xa, v_split, nB, eqB);
            SWIRL_LAUNCH_CHECK(ctx);
        }
        auto suffix = [](const uint32_t* base, int n_vars, int t) { return base + ((size_t(2) << n_vars) - (size_t(2) << (n_vars - t))) * 4; };

        RoundArgs a{};
        a.tree = layer_ptr(round + 1);
        a.partials = rs->d_partials;
        a.ticket = rs->d_ticket;
        a.result = rs->d_result;
        for (int k = 0; k < 4; k++) a.lambda[k] = lambda.c[k];
        for (int k = 0; k < 4; k++) a.c[k] = cst[round + 1].c[k];
        const Ext tail_unit = ext_mul(lambda, ext_mul(cst[round + 1], cst[round + 1]));  // inner() of a constant row
        const size_t rows_tree = (S[round + 1] + 1) / 2;  // stored rows of this layer's table
        std::vector<Ext> rho;
        Ext e_bound = one;
        // what this layer's sumcheck proves (verifier/fractional_sumcheck_gkr.rs:100-104): p(mu) + lambda q(mu) of the layer above
        const Ext mu_prev = xi_prev[0];
        Ext claim = ext_add(ext_lerp(prev_claims[0], prev_claims[2], mu_prev), ext_mul(lambda, ext_lerp(prev_claims[1], prev_claims[3], mu_prev)));
        size_t rows = rows_tree;  // stored rows of the table the next kernel reads
        int cur = 0;
        // Host side of a small table: stored rows (p0, q0, p1, q1); rows beyond are the constant (0, c, 0, c).
        using HRow = std::array<Ext, 4>;
        const Ext c_tail = cst[round + 1];
        const HRow const_row{bb::ext_zero(), c_tail, bb::ext_zero(), c_tail};
        std::vector<HRow> ht;
        auto row_at = [&](const std::vector<HRow>& t, size_t i) -> const HRow& { return i < t.size() ? t[i] : const_row; };
        const int sr_host = std::max(0, round - GKR_HOST_LOG);  // first round that runs on the host
        if (sr_host == 0) {  // the whole layer: rows straight from the tree copy
            const std::vector<uint32_t>& lay = top_layers[round + 1];
            ht.resize(rows_tree);
            for (size_t x = 0; x < rows_tree; x++)
                for (int k = 0; k < 4; k++) {
                    const size_t node = 2 * x + (k >> 1);
                    ht[x][k] = node < S[round + 1] ? ext_from_words(&lay[node * 8 + 4 * (k & 1)]) : const_row[k];
                }
        }
        // s(X) = e_bound * eq1(xi_sr, X) * t(X): t(1) and the leading coefficient come from the table sweep (the tail adds a
        // constant), t(0) from s(0) = claim - s(1), then t(2), t(3) by extrapolating the quadratic
        auto finish_round = [&](int sr, const Ext& t1_swept, const Ext& c2, size_t y_tail) -> int {
            const int m = round - 1 - sr;
            // the constant tail's share of t(X): lambda c^2 * sum_{y >= y_tail} eq(xi[sr+1..round), y)
            const Ext tail = ext_mul(tail_unit, ext_sub(one, eq_prefix_sum(xi_prev, sr + 1, m, y_tail)));
            uint32_t* out = h_polys + (poly_off + sr) * 12;
            const Ext xs = xi_prev[sr], one_minus_xs = ext_one_minus(xs);
            const Ext t1 = ext_add(t1_swept, tail);
            const Ext s1 = ext_mul(ext_mul(e_bound, xs), t1);
            const Ext den = ext_mul(e_bound, one_minus_xs);
            if (hp::is_zero(den)) {
                set_error("degenerate sumcheck challenge (eq factor is zero)");
                return SWIRL_ERR_INVALID;
            }
            const Ext s0 = ext_sub(claim, s1);
            const Ext t0 = ext_mul(s0, bb::ext_inv(den));
            const Ext c2x2 = ext_add(c2, c2);
            const Ext t2 = ext_add(ext_sub(ext_add(t1, t1), t0), c2x2);
            const Ext t3 = ext_add(ext_sub(ext_add(ext_add(t1, t1), t1), ext_add(t0, t0)), ext_add(ext_add(c2x2, c2x2), c2x2));
            // eq1(xs, 2) = 3 xs - 1, eq1(xs, 3) = 5 xs - 2
            const Ext xs2 = ext_add(xs, xs), xs3 = ext_add(xs2, xs), xs5 = ext_add(xs3, xs2);
            const Ext s2 = ext_mul(ext_mul(e_bound, ext_sub(xs3, one)), t2);
            const Ext s3 = ext_mul(ext_mul(e_bound, ext_sub(xs5, ext_add(one, one))), t3);
            const Ext sX[3] = {s1, s2, s3};
            for (int X = 0; X < 3; X++) {
                memcpy(out + 4 * X, sX[X].c, 16);
                tr.observe_ext(sX[X]);
            }
            const Ext r = tr.sample_ext();
            rho.push_back(r);
            {  // claim <- s(r): cubic through (0, s0), (1, s1), (2, s2), (3, s3)
                const Ext r1 = ext_sub(r, one), r2 = ext_sub(r1, one), r3 = ext_sub(r2, one);
                const Ext l0 = bb::ext_neg(ext_mul_base(ext_mul(ext_mul(r1, r2), r3), inv6));
                const Ext l1 = ext_mul_base(ext_mul(ext_mul(r, r2), r3), inv2);
                const Ext l2 = bb::ext_neg(ext_mul_base(ext_mul(ext_mul(r, r1), r3), inv2));
                const Ext l3 = ext_mul_base(ext_mul(ext_mul(r, r1), r2), inv6);
                claim = ext_add(ext_add(ext_mul(l0, s0), ext_mul(l1, s1)), ext_add(ext_mul(l2, s2), ext_mul(l3, s3)));
            }
            e_bound = ext_mul(e_bound, hp::eq1(xi_prev[sr], r));
            for (int k = 0; k < 4; k++) a.r[k] = r.c[k];
            return 0;
        };
        // ---- device rounds: launch (all up front when the round link is on), then one exchange per round ------------
        struct Launched {
            uint32_t seq;
            size_t y_tail, ny;
            bool exports;
        };
        std::vector<Launched> launched;
        uint32_t res_words[8];
        auto launch_round = [&](int sr, bool linked) -> int {
            // tables of the remaining variables [sr+1, round)
            if (sr + 1 < v_split) {
                a.a_bits = v_split - 1 - sr;
                a.A = suffix(eqA, nA, sr);
                a.B = suffix(eqB, nB, 0);
            } else {
                a.a_bits = 0;
                a.A = nullptr;
                a.B = suffix(eqB, nB, sr + 1 - v_split);
            }
            size_t y_tail;
            a.last_rows = sr == sr_host - 1 ? rs->d_result + 64 : nullptr;  // the table the host continues from
            a.link = linked ? link_make(rs, sr > 0) : RoundLink{};
            auto set_runs = [&]() {  // runs of >= 2^10 pairs, at least four per block of the grid; otherwise the plain sweep
                a.g_bits = 0;
                if (a.a_bits < 10) return;
                const size_t want = (size_t)4 * round_grid(ctx, a.ny, 3);
                int gb = a.a_bits;
                while (gb > 10 && (a.ny >> gb) < want) gb--;
                if ((a.ny >> gb) >= want) a.g_bits = gb;
            };
            {
                SwirlTimed timed(ctx, SWIRL_T_GKR);
                if (sr == 0) {
                    a.rows_in = rows_tree;
                    a.ny = y_tail = (rows_tree + 1) / 2;
                    set_runs();
                    if (a.g_bits)
                        gkr_round_kernel<true, false, true><<<round_grid(ctx, a.ny, 3), GKR_BLOCK, 0, ctx->stream>>>(a);
                    else
                        gkr_round_kernel<true, false, false><<<round_grid(ctx, a.ny), GKR_BLOCK, 0, ctx->stream>>>(a);
                } else if (sr == 1) {
                    a.rows_in = rows_tree;
                    a.ny = y_tail = (rows_tree + 3) / 4;
                    set_runs();
                    a.out = tab[0];
                    a.out_stride = tab_stride[0];
                    if (a.g_bits)
                        gkr_round_kernel<true, true, true><<<round_grid(ctx, a.ny, 3), GKR_BLOCK, 0, ctx->stream>>>(a);
                    else
                        gkr_round_kernel<true, true, false><<<round_grid(ctx, a.ny, 3), GKR_BLOCK, 0, ctx->stream>>>(a);
                    rows = 2 * a.ny;
                    cur = 0;
                } else {
                    a.in = tab[cur];
                    a.in_stride = tab_stride[cur];
                    a.rows_in = rows;
                    a.ny = y_tail = (rows + 3) / 4;
                    set_runs();
                    a.out = tab[cur ^ 1];
                    a.out_stride = tab_stride[cur ^ 1];
                    if (a.g_bits)
                        gkr_round_kernel<false, true, true><<<round_grid(ctx, a.ny, 3), GKR_BLOCK, 0, ctx->stream>>>(a);
                    else
                        gkr_round_kernel<false, true, false><<<round_grid(ctx, a.ny, 3), GKR_BLOCK, 0, ctx->stream>>>(a);
                    rows = 2 * a.ny;
                    cur ^= 1;
                }
                SWIRL_LAUNCH_CHECK(ctx);
            }
            launched.push_back({a.link.seq, y_tail, a.ny, a.last_rows != nullptr});
            return 0;
        };
        // Linked rounds are enqueued ONE ahead: kernel sr + 1 is launched right after kernel sr got its challenge, so
        // no CUDA call is ever made while an enqueued kernel waits for a mail that has not been sent (ext.cuh).
        const bool linked = ctx->round_link;
        if (linked && sr_host > 0) {
            link_begin(ctx, rs, 0, 8);
            rc = launch_round(0, true);
        }
        for (int sr = 0; sr < round && rc == 0; sr++) {
            if (sr >= sr_host) {
                // ---- host round: fold the small table with the previous challenge, then sweep it -------------------
                if (sr > 0) {
                    const Ext r_prev = rho.back();
                    const size_t ny = (ht.size() + 3) / 4;
                    std::vector<HRow> nt(2 * ny);
                    for (size_t y = 0; y < ny; y++)
                        for (int k = 0; k < 4; k++) {
                            nt[2 * y][k] = ext_lerp(row_at(ht, 4 * y)[k], row_at(ht, 4 * y + 1)[k], r_prev);
                            nt[2 * y + 1][k] = ext_lerp(row_at(ht, 4 * y + 2)[k], row_at(ht, 4 * y + 3)[k], r_prev);
                        }
                    ht.swap(nt);
                }
                const size_t ny = (ht.size() + 1) / 2;
                const int m = round - 1 - sr;
                Ext sw[2] = {bb::ext_zero(), bb::ext_zero()};
                for (size_t y = 0; y < ny; y++) {
                    Ext E = one;  // eq(xi[sr+1 .. round), y)
                    for (int b = 0; b < m; b++) E = ext_mul(E, hp::eq1(xi_prev[sr + 1 + b], ((y >> b) & 1) != 0));
                    Ext lo[4], hi[4];
                    for (int k = 0; k < 4; k++) {
                        lo[k] = row_at(ht, 2 * y)[k];
                        hi[k] = row_at(ht, 2 * y + 1)[k];
                    }
                    accumulate(lo, hi, lambda, E, sw);
                }
                rc = finish_round(sr, sw[0], sw[1], ny);
                continue;
            }
            // ---- device round ---------------------------------------------------------------------------------------
            if (linked) {
                if (sr > 0) link_send(rs, launched[sr].seq, rho.back());
                if (sr + 1 < sr_host) rc = launch_round(sr + 1, true);
                if (rc == 0) rc = link_recv(ctx, rs, launched[sr].seq, 0, 8, res_words);
                if (rc != 0) break;
            } else {
                rc = launch_round(sr, false);
                if (rc != 0) break;
                SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
                memcpy(res_words, rs->h_result, 32);
            }
            const Launched& L = launched[sr];
            if (L.exports) {  // this round's table (2 ny <= 2^(GKR_HOST_LOG+1) rows) for the host rounds that follow
                ht.resize(2 * L.ny);
                for (size_t i = 0; i < ht.size(); i++)
                    for (int k = 0; k < 4; k++) ht[i][k] = ext_from_words(rs->h_result + 64 + i * 16 + 4 * k);
            }
            rc = finish_round(sr, ext_from_words(res_words), ext_from_words(res_words + 4), L.y_tail);
        }
        if (rc != 0 && linked) {  // release the kernels that still wait for a challenge
            link_abort(rs);
            cudaStreamSynchronize(ctx->stream);
        }
        if (rc != 0) break;
        // claims: fold the final 2-row table with the last challenge
        {
            const Ext r_last = rho.back();
            for (int k = 0; k < 4; k++) {
                const Ext v = ext_lerp(row_at(ht, 0)[k], row_at(ht, 1)[k], r_last);
                memcpy(h_claims + (size_t)round * 16 + 4 * k, v.c, 16);
            }
        }
        uint32_t* cl = h_claims + (size_t)round * 16;
        for (int i = 0; i < 4; i++) {
            prev_claims[i] = ext_from_words(cl + 4 * i);
            tr.observe_ext(prev_claims[i]);
        }
        const Ext mu = tr.sample_ext();
        xi_prev.assign(1, mu);
        xi_prev.insert(xi_prev.end(), rho.begin(), rho.end());
        poly_off += round;
    }
    for (int b = 0; b < n; b++) memcpy(h_xi + 4 * b, xi_prev[b].c, 16);
    return rc;
}

extern "C" int swirl_gkr_fractional_sumcheck(swirl_ctx* ctx, swirl_transcript* ts, const uint32_t* d_leaves, int log_n,
                                             int assert_zero, uint32_t h_frac_sum[8], uint32_t* h_claims,
                                             uint32_t* h_polys, uint32_t* h_xi) {
    SWIRL_REQUIRE(log_n >= 1 && log_n <= 27, "log_n must be in [1, 27]");
    return swirl_gkr_fractional_sumcheck_padded(ctx, ts, d_leaves, uint64_t(1) << log_n, nullptr, log_n, assert_zero,
                                                h_frac_sum, h_claims, h_polys, h_xi);
}
