// WHIR opening proof of the stacked PCS.
//
// Replaces (reference, relative to /root/reference):
//   crates/cuda-backend/src/whir.rs:63-560             prove_whir_opening_gpu
//   crates/cuda-backend/cuda/src/whir.cu               batching / sumcheck / fold / w-accumulate kernels
// Semantics = crates/stark-backend/src/prover/whir.rs:78-352 (prove_whir_opening).
//
// Design: mu-batching happens BEFORE the basis changes (they are F-linear, so
// sum_j mu^j T(col_j) = T(sum_j mu^j col_j)): one coalesced sweep over the H x W stacked matrices
// leaves a single EF column, on whose 4 coordinate columns the chunk iDFT + zeta (ntt.cu) and the
// full zeta transform (mle.cu) run.  Each sumcheck round is one kernel that folds with the
// previous challenge on the fly, writes the half-size tables and accumulates s(1), s(2) (grid-wide
// reduction into mapped pinned memory).  Round codewords are RS-encoded as 4 base columns with the
// batched NTT and committed with the Merkle kernels of the commit path; opened rows / Merkle
// paths are gathered on the device.  The host side only runs the transcript.
#include <cstring>
#include <vector>

#include "ext.cuh"
#include "kernels.cuh"
#include "pcs.cuh"
#include "transcript.hpp"

namespace swirl {

using bb::ext_add;
using bb::ext_mul;
using bb::ext_sub;

constexpr int WH_BLOCK = 256;

// acc[i] (+)= sum_c mu[c] * M[c*H + i]; output as 4 coordinate columns (stride H)
__global__ void __launch_bounds__(WH_BLOCK)
whir_batch_kernel(const uint32_t* __restrict__ M, size_t H, uint32_t W, const uint32_t* __restrict__ mu_pows,
                  uint32_t* __restrict__ acc, int accumulate) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H) return;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (accumulate) {
        a0 = acc[i]; a1 = acc[H + i]; a2 = acc[2 * H + i]; a3 = acc[3 * H + i];
    }
    const uint32_t* p = M + i;
    // four columns per Montgomery reduction: sum of four 62-bit products < 2^64 (bb::dot4)
    uint32_t c = 0;
    for (; c + 4 <= W; c += 4) {
        uint32_t x[4];
        uint4 m[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            x[j] = __ldg(p + (size_t)(c + j) * H);
            m[j] = __ldg(reinterpret_cast<const uint4*>(mu_pows) + c + j);
        }
        a0 = bb::add(a0, bb::dot4(m[0].x, x[0], m[1].x, x[1], m[2].x, x[2], m[3].x, x[3]));
        a1 = bb::add(a1, bb::dot4(m[0].y, x[0], m[1].y, x[1], m[2].y, x[2], m[3].y, x[3]));
        a2 = bb::add(a2, bb::dot4(m[0].z, x[0], m[1].z, x[1], m[2].z, x[2], m[3].z, x[3]));
        a3 = bb::add(a3, bb::dot4(m[0].w, x[0], m[1].w, x[1], m[2].w, x[2], m[3].w, x[3]));
    }
    for (; c < W; c++) {
        const uint32_t x = __ldg(p + (size_t)c * H);
        const uint4 m = __ldg(reinterpret_cast<const uint4*>(mu_pows) + c);
        a0 = bb::add(a0, bb::mul(m.x, x));
        a1 = bb::add(a1, bb::mul(m.y, x));
        a2 = bb::add(a2, bb::mul(m.z, x));
        a3 = bb::add(a3, bb::mul(m.w, x));
    }
    acc[i] = a0; acc[H + i] = a1; acc[2 * H + i] = a2; acc[3 * H + i] = a3;
}

struct WhirRoundArgs {
    const uint32_t *f_in, *w_in;  // EF arrays of `n` entries
    uint32_t *f_out, *w_out;      // FOLD: n/2 entries
    size_t n;
    uint32_t alpha[4];  // FOLD: previous challenge
    uint32_t* partials;
    unsigned int* ticket;
    uint32_t* result;  // 8 words: s(1), s(2)
    RoundLink link;    // seq != 0: alpha arrives through the mailbox, the result words carry the ready mark (ext.cuh)
};

// MODE 0: s from the table as it is.  MODE 1: fold pairs with alpha, write, and accumulate s of
// the folded table.  MODE 2: fold and write only.
template <int MODE>
__global__ void __launch_bounds__(WH_BLOCK) whir_round_kernel(WhirRoundArgs a) {
    Ext alpha = Ext{{a.alpha[0], a.alpha[1], a.alpha[2], a.alpha[3]}};
    if (MODE != 0 && !link_wait(a.link, alpha)) return;
    Ext s1 = bb::ext_zero(), s2 = bb::ext_zero();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 2) {
        for (size_t y = tid; y < (a.n >> 1); y += stride) {
            st_ext(a.f_out + y * 4, ext_lerp(ldg_ext(a.f_in + 8 * y), ldg_ext(a.f_in + 8 * y + 4), alpha));
            st_ext(a.w_out + y * 4, ext_lerp(ldg_ext(a.w_in + 8 * y), ldg_ext(a.w_in + 8 * y + 4), alpha));
        }
        return;
    }
    const size_t ny = MODE == 1 ? a.n >> 2 : a.n >> 1;
    for (size_t y = tid; y < ny; y += stride) {
        Ext f0, f1, w0, w1;
        if (MODE == 1) {
            f0 = ext_lerp(ldg_ext(a.f_in + 16 * y), ldg_ext(a.f_in + 16 * y + 4), alpha);
            f1 = ext_lerp(ldg_ext(a.f_in + 16 * y + 8), ldg_ext(a.f_in + 16 * y + 12), alpha);
            w0 = ext_lerp(ldg_ext(a.w_in + 16 * y), ldg_ext(a.w_in + 16 * y + 4), alpha);
            w1 = ext_lerp(ldg_ext(a.w_in + 16 * y + 8), ldg_ext(a.w_in + 16 * y + 12), alpha);
            st_ext(a.f_out + 8 * y, f0);
            st_ext(a.f_out + 8 * y + 4, f1);
            st_ext(a.w_out + 8 * y, w0);
            st_ext(a.w_out + 8 * y + 4, w1);
        } else {
            f0 = ldg_ext(a.f_in + 8 * y);
            f1 = ldg_ext(a.f_in + 8 * y + 4);
            w0 = ldg_ext(a.w_in + 8 * y);
            w1 = ldg_ext(a.w_in + 8 * y + 4);
        }
        // X = 1: (f1, w1);  X = 2: (2 f1 - f0, 2 w1 - w0)
        s1 = ext_add(s1, ext_mul(f1, w1));
        s2 = ext_add(s2, ext_mul(ext_sub(ext_add(f1, f1), f0), ext_sub(ext_add(w1, w1), w0)));
    }
    uint32_t v[8] = {s1.c[0], s1.c[1], s1.c[2], s1.c[3], s2.c[0], s2.c[1], s2.c[2], s2.c[3]};
    grid_sum<8>(v, a.partials, a.ticket, a.result, link_result_tag(a.link.seq));
}

struct PowArgs {
    uint32_t p[28][4];  // z^(2^b)
};

// sum_i g_i * z^i over the 4 coordinate columns of g (mle eval at (z, z^2, z^4, ...), whir.rs:221-223)
__global__ void __launch_bounds__(WH_BLOCK)
whir_ood_kernel(const uint32_t* __restrict__ g_soa, size_t n, size_t col_stride, PowArgs zp, int dim,
                uint32_t* partials, unsigned int* ticket, uint32_t* result) {
    Ext acc = bb::ext_zero();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        Ext term = Ext{{g_soa[i], g_soa[col_stride + i], g_soa[2 * col_stride + i], g_soa[3 * col_stride + i]}};
        for (int b = 0; b < dim; b++)
            if ((i >> b) & 1) term = ext_mul(term, Ext{{zp.p[b][0], zp.p[b][1], zp.p[b][2], zp.p[b][3]}});
        acc = ext_add(acc, term);
    }
    uint32_t v[4] = {acc.c[0], acc.c[1], acc.c[2], acc.c[3]};
    grid_sum<4>(v, partials, ticket, result);
}

// w[x] += gamma * eq(x, pow(z0)) + sum_q gamma^(q+2) * eq(x, pow(z_q)),  z_q in F (whir.rs:310-325).
// eq(x, pow(z)) = A_z[x_lo] * B_z[x_hi] (product over the low / high index bits), so the per-query
// half tables are built first (whir_eq_halves_kernel) and an entry costs one base multiply and one
// EF x F multiply per query instead of 2 * dim multiplies.
// tabs: per query [A (2^lo_bits) | B (2^hi_bits)] base-field words.
__global__ void __launch_bounds__(WH_BLOCK)
whir_eq_halves_kernel(const uint32_t* __restrict__ zs, int nq, int lo_bits, int hi_bits, uint32_t* __restrict__ tabs) {
    const size_t per_q = (size_t(1) << lo_bits) + (size_t(1) << hi_bits);
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= per_q * (size_t)nq) return;
    const size_t q = g / per_q, i = g % per_q;
    const bool hi = i >= (size_t(1) << lo_bits);
    const size_t idx = hi ? i - (size_t(1) << lo_bits) : i;
    const int bits = hi ? hi_bits : lo_bits;
    uint32_t zp = __ldg(zs + q), e = bb::R1;
    if (hi)
        for (int b = 0; b < lo_bits; b++) zp = bb::sqr(zp);  // z^(2^lo_bits)
    for (int b = 0; b < bits; b++) {
        e = bb::mul(e, ((idx >> b) & 1) ? zp : bb::sub(bb::R1, zp));
        zp = bb::sqr(zp);
    }
    tabs[g] = e;
}
__global__ void __launch_bounds__(WH_BLOCK)
whir_w_accumulate_kernel(uint32_t* __restrict__ w, size_t n, int dim, PowArgs z0p, const uint32_t* __restrict__ tabs,
                         int lo_bits, const uint32_t* __restrict__ gamma_pows /* [0] = gamma, [1+q] = gamma^(q+2) */, int nq) {
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    Ext e0 = bb::ext_one();
    for (int b = 0; b < dim; b++) {
        const Ext zb = Ext{{z0p.p[b][0], z0p.p[b][1], z0p.p[b][2], z0p.p[b][3]}};
        e0 = ext_mul(e0, ((x >> b) & 1) ? zb : ext_one_minus(zb));
    }
    Ext acc = ext_add(ld_ext(w + 4 * x), ext_mul(ldg_ext(gamma_pows), e0));
    const size_t lo = x & ((size_t(1) << lo_bits) - 1), hi = x >> lo_bits;
    const size_t per_q = (size_t(1) << lo_bits) + (size_t(1) << (dim - lo_bits));
    for (int q = 0; q < nq; q++) {
        const uint32_t* t = tabs + (size_t)q * per_q;
        const uint32_t e = bb::mul(__ldg(t + lo), __ldg(t + (size_t(1) << lo_bits) + hi));
        acc = ext_add(acc, bb::ext_mul_base(ldg_ext(gamma_pows + 4 * (q + 1)), e));
    }
    st_ext(w + 4 * x, acc);
}

static int wh_grid(const swirl_ctx* ctx, size_t items) {
    size_t blocks = (items + WH_BLOCK - 1) / WH_BLOCK;
    const size_t cap = (size_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    return blocks ? (int)blocks : 1;
}

static size_t whir_words(int m, int log_blowup, const swirl_whir_config* cfg, size_t n_commits, const uint64_t* widths) {
    const int k = cfg->k, R = cfg->num_rounds;
    size_t n = 1 + (size_t)R * k * 8 + (size_t)(R - 1) * 12 + (size_t)R * k + R;
    for (size_t i = 0; i < n_commits; i++)
        n += (size_t)cfg->num_queries[0] * ((widths[i] << k) + (size_t)(m + log_blowup - k) * 8);
    for (int r = 1; r < R; r++) n += (size_t)cfg->num_queries[r] * ((size_t(4) << k) + (size_t)(m + log_blowup - r - k) * 8);
    n += size_t(4) << (m - R * k);
    return n;
}

static int check_cfg(const swirl_pcs_params* p, const swirl_whir_config* cfg) {
    SWIRL_REQUIRE(p && cfg, "null argument");
    SWIRL_REQUIRE(cfg->k >= 1 && cfg->k == p->k_whir, "whir k must equal the commitment's k_whir");
    SWIRL_REQUIRE(cfg->num_rounds >= 1 && cfg->num_rounds <= 32, "num_rounds");
    SWIRL_REQUIRE(cfg->num_rounds * cfg->k <= p->l_skip + p->n_stack, "more sumcheck rounds than variables");
    SWIRL_REQUIRE(p->l_skip + p->n_stack + p->log_blowup - (cfg->num_rounds - 1) - cfg->k >= 0, "domain too small");
    for (int r = 0; r < cfg->num_rounds; r++) SWIRL_REQUIRE(cfg->num_queries[r] >= 0, "num_queries");
    return 0;
}

}  // namespace swirl

using namespace swirl;

extern "C" size_t swirl_whir_proof_words(const swirl_pcs_params* params, const swirl_whir_config* cfg, size_t n_commits,
                                         const uint64_t* widths) {
    if (check_cfg(params, cfg) != 0 || (!widths && n_commits)) return 0;
    return whir_words(params->l_skip + params->n_stack, params->log_blowup, cfg, n_commits, widths);
}

extern "C" int swirl_whir_open(swirl_ctx* ctx, swirl_transcript* ts, const swirl_whir_config* cfg,
                               const swirl_pcs* const* pcs, size_t n_commits, const uint32_t* h_u, uint32_t* h_proof,
                               size_t proof_words) {
    SWIRL_REQUIRE(ctx && ts && cfg && pcs && n_commits >= 1 && h_u && h_proof, "null argument");
    const swirl_pcs_params params = pcs[0]->params;
    SWIRL_TRY(check_cfg(&params, cfg));
    SWIRL_CUDA(cudaSetDevice(ctx->device));
    const int l_skip = params.l_skip, log_blowup = params.log_blowup, k = cfg->k, R = cfg->num_rounds;
    int m = l_skip + params.n_stack;
    const size_t H = size_t(1) << m;
    std::vector<uint64_t> widths(n_commits);
    size_t total_w = 0;
    for (size_t i = 0; i < n_commits; i++) {
        SWIRL_REQUIRE(pcs[i] && pcs[i]->layout.height == H && pcs[i]->params.l_skip == l_skip &&
                          pcs[i]->params.log_blowup == log_blowup && pcs[i]->params.k_whir == k,
                      "commitments must share height and parameters");
        SWIRL_REQUIRE(pcs[i]->codeword_height == (H << log_blowup), "TreeHeightMismatch");
        widths[i] = pcs[i]->layout.width;
        total_w += widths[i];
    }
    SWIRL_REQUIRE(proof_words == whir_words(m, log_blowup, cfg, n_commits, widths.data()), "proof buffer size");
    Transcript tr(ts);
    RoundScratch* rs;
    SWIRL_TRY(round_scratch_get(ctx, &rs));

    // ---- section cursors of the flat proof ---------------------------------------------------
    uint32_t* p = h_proof;
    uint32_t* sec_mu = p; p += 1;
    uint32_t* sec_polys = p; p += (size_t)R * k * 8;
    uint32_t* sec_commits = p; p += (size_t)(R - 1) * 8;
    uint32_t* sec_ood = p; p += (size_t)(R - 1) * 4;
    uint32_t* sec_fold_pow = p; p += (size_t)R * k;
    uint32_t* sec_query_pow = p; p += R;
    std::vector<uint32_t*> sec_rows0(n_commits), sec_proofs0(n_commits);
    for (size_t i = 0; i < n_commits; i++) { sec_rows0[i] = p; p += (size_t)cfg->num_queries[0] * (widths[i] << k); }
    for (size_t i = 0; i < n_commits; i++) { sec_proofs0[i] = p; p += (size_t)cfg->num_queries[0] * (m + log_blowup - k) * 8; }
    std::vector<uint32_t*> sec_vals(R, nullptr), sec_proofs(R, nullptr);
    for (int r = 1; r < R; r++) { sec_vals[r] = p; p += (size_t)cfg->num_queries[r] * (size_t(4) << k); }
    for (int r = 1; r < R; r++) { sec_proofs[r] = p; p += (size_t)cfg->num_queries[r] * (m + log_blowup - r - k) * 8; }
    uint32_t* sec_final = p;

    auto t_prev = std::chrono::steady_clock::now();
    swirl::trace_mark(ctx, "whir", nullptr, &t_prev);
    // ---- mu batching -----------------------------------------------------------------------------
    uint32_t wm = 0;
    SWIRL_TRY(transcript_grind(ctx, ts, cfg->mu_pow_bits, &wm));
    sec_mu[0] = wm;
    const Ext mu = tr.sample_ext();
    std::vector<uint32_t> mu_pows(total_w * 4);
    {
        Ext a = bb::ext_one();
        for (size_t j = 0; j < total_w; j++) {
            memcpy(&mu_pows[4 * j], a.c, 16);
            a = ext_mul(a, mu);
        }
    }
    uint32_t *d_mu = nullptr, *soa = nullptr, *f[2] = {nullptr, nullptr}, *w[2] = {nullptr, nullptr};
    int max_q = 0;
    for (int r = 0; r < R; r++) max_q = cfg->num_queries[r] > max_q ? cfg->num_queries[r] : max_q;
    uint32_t *d_idx = nullptr, *d_zs = nullptr, *d_gam = nullptr, *d_open = nullptr;
    size_t open_words = 0;
    for (size_t i = 0; i < n_commits; i++) {
        const size_t wds = (size_t)cfg->num_queries[0] * ((widths[i] << k) + (size_t)(m + log_blowup - k) * 8);
        open_words = wds > open_words ? wds : open_words;
    }
    for (int r = 1; r < R; r++) {
        const size_t wds = (size_t)cfg->num_queries[r] * ((size_t(4) << k) + (size_t)(m + log_blowup - r - k) * 8);
        open_words = wds > open_words ? wds : open_words;
    }
    ArenaGuard scratch(ctx);  // the tables that live for the whole call: released on every return path
#define WH_ALLOC(ptr, count)                          \
    SWIRL_CUDA(dev_alloc(ctx, &(ptr), (count)));      \
    scratch.add(ptr)
    WH_ALLOC(d_mu, total_w * 4);
    WH_ALLOC(soa, H * 4);
    WH_ALLOC(f[0], H * 4);
    WH_ALLOC(f[1], H * 2);
    WH_ALLOC(w[0], H * 4);
    WH_ALLOC(w[1], H * 2);
    WH_ALLOC(d_idx, (size_t)max_q + 1);
    WH_ALLOC(d_zs, (size_t)max_q + 1);
    WH_ALLOC(d_gam, ((size_t)max_q + 1) * 4);
    WH_ALLOC(d_open, open_words + 4);
#undef WH_ALLOC
    SWIRL_CUDA(cudaMemcpyAsync(d_mu, mu_pows.data(), total_w * 16, cudaMemcpyHostToDevice, ctx->stream));
    {
        size_t off = 0;
        for (size_t i = 0; i < n_commits; i++) {
            whir_batch_kernel<<<(unsigned)((H + WH_BLOCK - 1) / WH_BLOCK), WH_BLOCK, 0, ctx->stream>>>(
                pcs[i]->stacked, H, (uint32_t)widths[i], d_mu + 4 * off, soa, i > 0 ? 1 : 0);
            SWIRL_LAUNCH_CHECK(ctx);
            off += widths[i];
        }
    }
    // RS message of the batched column (chunk iDFT + zeta over the l_skip bits), then MLE
    // coefficients -> hypercube evaluations over all m bits (whir.rs:124-133)
    if (l_skip > 0) SWIRL_TRY(chunk_coeffs(ctx, soa, H, soa, H, H, 4, l_skip));
    SWIRL_TRY(mle_zeta(ctx, soa, H, m, 4, false));
    SWIRL_TRY(ext_soa_to_aos(ctx, soa, f[0], H, H));
    // w = mobius_eq(u, .)
    {
        TensorArgs t;
        for (int b = 0; b < m; b++) {
            const Ext ub = Ext{{h_u[4 * b], h_u[4 * b + 1], h_u[4 * b + 2], h_u[4 * b + 3]}};
            const Ext w0 = ext_sub(bb::ext_one(), ext_add(ub, ub));
            memcpy(t.w0[b], w0.c, 16);
            memcpy(t.w1[b], ub.c, 16);
        }
        SWIRL_TRY(mle_tensor_table(ctx, t, m, w[0]));
    }

    swirl::trace_mark(ctx, "whir", "mu batch + tables", &t_prev);
    // ---- WHIR rounds ---------------------------------------------------------------------------
    int cur = 0;             // f[cur], w[cur] hold the current tables of n entries
    size_t n = H;
    uint32_t* rs_codeword = nullptr;  // previous round's codeword (N x 4) and its digest layers
    uint32_t* rs_layers = nullptr;
    size_t rs_height = 0;
    int log_rs = m + log_blowup;
    int rc = 0;
    size_t sc_i = 0;
    std::vector<uint32_t> h_idx(max_q + 1), h_zs(max_q + 1), h_gam(((size_t)max_q + 1) * 4);
    for (int wr = 0; wr < R && rc == 0; wr++) {
        const bool is_last = wr == R - 1;
        WhirRoundArgs a{};
        a.partials = rs->d_partials;
        a.ticket = rs->d_ticket;
        a.result = rs->d_result;
        // The k sumcheck rounds and the kernel that materialises the last fold.  With the round link they are all enqueued
        // first and take their challenge from the mailbox; the proof-of-work search between two rounds must then stay on
        // the host (a grind kernel would queue behind the kernels that wait for the challenge it precedes).
        const bool linked = ctx->round_link && cfg->folding_pow_bits <= 10 && k <= 32;
        uint32_t seqs[33] = {0};
        auto launch_round = [&](int round) -> int {  // round == k: fold and write only
            a.f_in = f[cur];
            a.w_in = w[cur];
            a.n = n;
            a.link = linked ? link_make(rs, round > 0) : RoundLink{};
            seqs[round] = a.link.seq;
            if (round == 0) {
                whir_round_kernel<0><<<wh_grid(ctx, n >> 1), WH_BLOCK, 0, ctx->stream>>>(a);
            } else {
                a.f_out = f[cur ^ 1];
                a.w_out = w[cur ^ 1];
                if (round < k)
                    whir_round_kernel<1><<<wh_grid(ctx, n >> 2), WH_BLOCK, 0, ctx->stream>>>(a);
                else
                    whir_round_kernel<2><<<wh_grid(ctx, n >> 1), WH_BLOCK, 0, ctx->stream>>>(a);
                cur ^= 1;
                n >>= 1;
            }
            SWIRL_LAUNCH_CHECK(ctx);
            return 0;
        };
        if (linked) {  // one kernel ahead of the exchange, see the rule in ext.cuh
            link_begin(ctx, rs, 0, 8);
            SWIRL_TRY(launch_round(0));
        }
        for (int round = 0; round < k; round++, sc_i++) {
            uint32_t* s = sec_polys + sc_i * 8;
            if (linked) {
                rc = launch_round(round + 1);  // kernel `round` has its challenge already
                if (rc == 0) rc = link_recv(ctx, rs, seqs[round], 0, 8, s);
                if (rc != 0) {
                    link_abort(rs);
                    cudaStreamSynchronize(ctx->stream);
                    break;
                }
            } else {
                SWIRL_TRY(launch_round(round));
                SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
                memcpy(s, rs->h_result, 32);
            }
            tr.observe_ext(Ext{{s[0], s[1], s[2], s[3]}});
            tr.observe_ext(Ext{{s[4], s[5], s[6], s[7]}});
            rc = transcript_grind(ctx, ts, cfg->folding_pow_bits, &sec_fold_pow[sc_i]);
            if (rc != 0) {
                if (linked) {
                    link_abort(rs);
                    cudaStreamSynchronize(ctx->stream);
                }
                break;
            }
            const Ext alpha = tr.sample_ext();
            memcpy(a.alpha, alpha.c, 16);
            if (linked) link_send(rs, seqs[round + 1], alpha);
        }
        if (rc != 0) break;
        swirl::trace_mark(ctx, "whir", "sumcheck rounds", &t_prev);
        // materialise the last fold of this WHIR round
        if (!linked) SWIRL_TRY(launch_round(k));
        // g = MLE coefficients of f (4 coordinate columns)
        SWIRL_TRY(ext_aos_to_soa(ctx, f[cur], soa, n, n));
        SWIRL_TRY(mle_zeta(ctx, soa, n, m - k, 4, true));
        uint32_t *g_codeword = nullptr, *g_layers = nullptr;
        Ext z0 = bb::ext_zero();
        PowArgs z0p{};
        if (!is_last) {
            const int log_N = log_rs - 1;
            const size_t N = size_t(1) << log_N;
            SWIRL_CUDA(dev_alloc(ctx, &g_codeword, N * 4));
            const size_t S = N >> k;
            SWIRL_CUDA(dev_alloc(ctx, &g_layers, (2 * S) * 8));
            SWIRL_TRY(rs_encode(ctx, soa, n, n, 4, 0, log_N - (m - k), g_codeword));
            SWIRL_TRY(merkle_commit(ctx, g_codeword, N, 4, k, g_layers));
            uint32_t* root = sec_commits + 8 * wr;
            SWIRL_CUDA(cudaMemcpyAsync(root, g_layers + (2 * S - 2) * 8, 32, cudaMemcpyDeviceToHost, ctx->stream));
            if (linked) SWIRL_CUDA(link_flag_fetch(ctx, rs));
            SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
            if (linked && link_aborted(rs)) {
                set_error("round link: the fold kernel gave up waiting for its challenge");
                rc = SWIRL_ERR_INVALID;
                break;
            }
            tr.observe_digest(root);
            z0 = tr.sample_ext();
            Ext zp = z0;
            for (int b = 0; b < m - k; b++) {
                memcpy(z0p.p[b], zp.c, 16);
                zp = bb::ext_sqr(zp);
            }
            whir_ood_kernel<<<wh_grid(ctx, n), WH_BLOCK, 0, ctx->stream>>>(soa, n, n, z0p, m - k, rs->d_partials, rs->d_ticket,
                                                                          rs->d_result);
            SWIRL_LAUNCH_CHECK(ctx);
            SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
            uint32_t* y0 = sec_ood + 4 * wr;
            memcpy(y0, rs->h_result, 16);
            tr.observe_ext(Ext{{y0[0], y0[1], y0[2], y0[3]}});
        } else {
            // final polynomial: coefficients to the host (interleave the 4 coordinate columns)
            std::vector<uint32_t> cols(n * 4);
            SWIRL_CUDA(cudaMemcpyAsync(cols.data(), soa, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
            if (linked) SWIRL_CUDA(link_flag_fetch(ctx, rs));
            SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
            if (linked && link_aborted(rs)) {
                set_error("round link: the fold kernel gave up waiting for its challenge");
                rc = SWIRL_ERR_INVALID;
                break;
            }
            for (size_t i = 0; i < n; i++) {
                for (int c = 0; c < 4; c++) sec_final[4 * i + c] = cols[c * n + i];
                tr.observe_ext(Ext{{sec_final[4 * i], sec_final[4 * i + 1], sec_final[4 * i + 2], sec_final[4 * i + 3]}});
            }
        }
        swirl::trace_mark(ctx, "whir", "fold+commit+ood", &t_prev);
        // ---- query phase ---------------------------------------------------------------------------
        const int nq = cfg->num_queries[wr];
        SWIRL_TRY(transcript_grind(ctx, ts, cfg->query_phase_pow_bits, &sec_query_pow[wr]));
        swirl::trace_mark(ctx, "whir", "  query grind", &t_prev);
        const uint32_t omega = bb::two_adic_generator(log_rs - k);
        for (int q = 0; q < nq; q++) {
            h_idx[q] = tr.sample_bits(log_rs - k);
            h_zs[q] = bb::pow(omega, h_idx[q]);
        }
        swirl::trace_mark(ctx, "whir", "  sample indices", &t_prev);
        const size_t depth = (size_t)(log_rs - k);
        if (nq > 0) {
            SWIRL_CUDA(cudaMemcpyAsync(d_idx, h_idx.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, ctx->stream));
            if (wr == 0) {
                for (size_t ci = 0; ci < n_commits; ci++) {
                    const size_t row_words = (size_t)nq * (widths[ci] << k), path_words = (size_t)nq * depth * 8;
                    if (pcs[ci]->open_fn) {
                        // the tree lives elsewhere (sharded commitment): the owner of each query supplies rows and path
                        SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
                        const int orc = pcs[ci]->open_fn(pcs[ci]->open_user, h_idx.data(), (size_t)nq, d_open, d_open + row_words);
                        if (orc != 0) {
                            set_error("external opening callback failed");
                            return orc;
                        }
                    } else {
                        SWIRL_TRY(pcs_open_rows(ctx, pcs[ci], k, d_idx, nq, d_open));  // cached codeword, or re-encoded by column groups
                        SWIRL_TRY(merkle_query_proofs(ctx, pcs[ci]->layers, pcs[ci]->query_stride, d_idx, nq, d_open + row_words));
                    }
                    swirl::trace_mark(ctx, "whir", "  open kernels", &t_prev);
                    SWIRL_CUDA(swirl::d2h_staged(ctx, sec_rows0[ci], d_open, row_words * 4));
                    SWIRL_CUDA(swirl::d2h_staged(ctx, sec_proofs0[ci], d_open + row_words, path_words * 4));
                }
            } else {
                SWIRL_REQUIRE(rs_codeword && rs_layers, "RsTreeNone");
                const size_t row_words = (size_t)nq * (size_t(4) << k), path_words = (size_t)nq * depth * 8;
                SWIRL_TRY(matrix_open_rows(ctx, rs_codeword, rs_height, 4, rs_height >> k, k, d_idx, nq, d_open));
                SWIRL_TRY(merkle_query_proofs(ctx, rs_layers, rs_height >> k, d_idx, nq, d_open + row_words));
                SWIRL_CUDA(swirl::d2h_staged(ctx, sec_vals[wr], d_open, row_words * 4));
                SWIRL_CUDA(swirl::d2h_staged(ctx, sec_proofs[wr], d_open + row_words, path_words * 4));
            }
        }
        swirl::trace_mark(ctx, "whir", "grind+queries", &t_prev);
        dev_free(ctx, rs_codeword);
        dev_free(ctx, rs_layers);
        rs_codeword = g_codeword;
        rs_layers = g_layers;
        rs_height = is_last ? 0 : size_t(1) << (log_rs - 1);
        const Ext gamma = tr.sample_ext();
        if (!is_last) {
            Ext gp = gamma;
            memcpy(&h_gam[0], gp.c, 16);
            gp = ext_mul(gp, gamma);
            for (int q = 0; q < nq; q++) {
                memcpy(&h_gam[4 * (q + 1)], gp.c, 16);
                gp = ext_mul(gp, gamma);
            }
            SWIRL_CUDA(cudaMemcpyAsync(d_gam, h_gam.data(), ((size_t)nq + 1) * 16, cudaMemcpyHostToDevice, ctx->stream));
            if (nq) SWIRL_CUDA(cudaMemcpyAsync(d_zs, h_zs.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, ctx->stream));
            {
                const int dim = m - k, lo_bits = dim / 2, hi_bits = dim - lo_bits;
                const size_t per_q = (size_t(1) << lo_bits) + (size_t(1) << hi_bits);
                uint32_t* tabs = nullptr;
                SWIRL_CUDA(dev_alloc(ctx, &tabs, per_q * (size_t)(nq ? nq : 1)));
                if (nq) {
                    whir_eq_halves_kernel<<<(unsigned)((per_q * nq + WH_BLOCK - 1) / WH_BLOCK), WH_BLOCK, 0, ctx->stream>>>(
                        d_zs, nq, lo_bits, hi_bits, tabs);
                    SWIRL_LAUNCH_CHECK(ctx);
                }
                whir_w_accumulate_kernel<<<(unsigned)((n + WH_BLOCK - 1) / WH_BLOCK), WH_BLOCK, 0, ctx->stream>>>(
                    w[cur], n, dim, z0p, tabs, lo_bits, d_gam, nq);
                SWIRL_LAUNCH_CHECK(ctx);
                dev_free(ctx, tabs);
            }
            // h_gam / h_zs are reused next round: make sure the copies have been consumed
            SWIRL_CUDA(swirl::stream_sync(ctx, __FILE__, __LINE__));
        }
        m -= k;
        log_rs -= 1;
    }
    dev_free(ctx, rs_codeword);  // the per-round codeword / digest layers rotate and are released as they are replaced
    dev_free(ctx, rs_layers);
    return rc;
}
